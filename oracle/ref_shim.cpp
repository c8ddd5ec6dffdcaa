// ref_shim.cpp — OUR extern "C" access layer over the unmodified reference (linked from
// oracle/_ref/libvvdec_ref.a).  TEST INFRASTRUCTURE ONLY: pins oracle/*.c and the flattener
// (vvdec_b200/vvdec_glue) against the reference's own functions, and provides the "reference"
// CPU baseline for bench.py.  Same construction trick as tests/vvdec_unit_test/vvdec_unit_test.cpp.
#include "ref_shim.h"
#include <mutex>
#include <memory>
#include <cstring>
#include <string>
#include <vector>
#include <sstream>
#include <list>
#include <map>
#include <array>
#include <atomic>
#include <condition_variable>
#include <thread>
#include <functional>
#include <algorithm>
#include <iostream>
// the reference keeps a few kernels reachable only through private members; the shim (test code) opens them
#define private public
#define protected public
#include "CommonLib/CommonDef.h"
#include "CommonLib/Rom.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/Slice.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/Picture.h"
#include "CommonLib/Quant.h"
#include "CommonLib/TrQuant.h"
#include "CommonLib/TrQuant_EMT.h"
#include "CommonLib/InterPrediction.h"
#include "CommonLib/InterpolationFilter.h"
#include "CommonLib/LoopFilter.h"
#include "CommonLib/SampleAdaptiveOffset.h"
#include "CommonLib/AdaptiveLoopFilter.h"
#include "../vvdec_b200/vvdec_glue/flatten_tu.h"

using namespace vvdec;

namespace vvdec { extern InvTrans* fastInvTrans[NUM_TRANS_TYPE][g_numTransformMatrixSizes]; }

static std::once_flag g_once;
static TCoeffOps g_scalarOps, g_simdOps;
static void globalInit()
{
  std::call_once( g_once, [] {
    initROM();
#if defined( TARGET_SIMD_X86 ) && ENABLE_SIMD_TCOEFF_OPS
    g_simdOps.initTCoeffOpsX86();
#endif
  } );
}

extern "C" const char* ref_simd_level( void )
{
  static std::string s;
#if defined( TARGET_SIMD_X86 )
  s = read_x86_extension_name();
#else
  s = "SCALAR";
#endif
  return s.c_str();
}

// ------------------------------------------------------------------------------------------------ K1
extern "C" void ref_dequant( int simd, int width, int maxX, int maxY, int scale, const int16_t* q, size_t qStride, int32_t* coef,
                             int rightShift, int inputMaximum, int32_t transformMaximum )
{
  globalInit();
  static Quant qs( nullptr, false ), qv( nullptr, true );
  ( simd ? qv : qs ).DeQuant( width, maxX, maxY, scale, q, qStride, coef, rightShift, inputMaximum, transformMaximum );
}

static InterPrediction* sharedIP() { static InterPrediction* ip = new InterPrediction(); return ip; }
static TrQuant*         sharedTQ() { static TrQuant* tq = new TrQuant( sharedIP() ); return tq; }

extern "C" void ref_inv_lfnst( int32_t* src, int32_t* dst, unsigned set, unsigned index, unsigned size, int zeroOutSize )
{
  globalInit();
  sharedTQ()->m_invLfnstNxN( src, dst, set, index, size, zeroOutSize );
}

extern "C" void ref_inv_1d( int simd, int trType, int n, const int32_t* src, int32_t* dst, int shift, int line, int skipLine,
                            int skipLine2, int clip, int32_t outMin, int32_t outMax )
{
  globalInit();
  g_tCoeffOps = simd ? g_simdOps : g_scalarOps;
  fastInvTrans[trType][getLog2( n ) - 1]( src, dst, shift, line, skipLine, skipLine2, clip != 0, outMin, outMax );
}

extern "C" void ref_cpy_resi_clip( int simd, const int32_t* src, int16_t* dst, ptrdiff_t stride, unsigned w, unsigned h,
                                   int32_t outMin, int32_t outMax, int32_t round, int32_t shift )
{
  globalInit();
  ( simd ? g_simdOps : g_scalarOps ).cpyResiClip[getLog2( w )]( src, dst, stride, w, h, outMin, outMax, round, shift );
}

// ---- fake parsed state for one CU == one TU -----------------------------------------------------
struct FakeCtx
{
  SPS       sps;
  PPS       pps;
  std::shared_ptr<PicHeader> ph = std::make_shared<PicHeader>();
  Slice     slice;
  CUChunkCache cuCache;
  TUChunkCache tuCache;
  CodingStructure cs{ &cuCache, &tuCache };
  CodingUnit cu;
};

extern "C" int ref_tu_case( const ref_tu_syntax* s, const int16_t* levels, int16_t* resi0, int16_t* resi1,
                            b200_tu* rec, int16_t* coefsOut, int32_t* numCoefs )
{
  globalInit();
  g_tCoeffOps = g_scalarOps;
  std::unique_ptr<FakeCtx> c( new FakeCtx );
  const ChromaFormat fmt = CHROMA_420;
  c->sps.setChromaFormatIdc( fmt );
  c->sps.setBitDepth( s->bitDepth );
  c->sps.setQpBDOffset( 6 * ( s->bitDepth - 8 ) );
  c->sps.setInternalMinusInputBitDepth( 0 );
  c->sps.setUseMTS( s->spsMTS );
  c->sps.setUseIntraMTS( s->spsIntraMTS );
  c->sps.setUseInterMTS( s->spsInterMTS );
  c->sps.setUseLFNST( s->spsLFNST );
  {
    ChromaQpMappingTableParams p;           // default-constructed: one table, one pivot (26 -> 26), slope 1 below / above
    p.m_qpBdOffset = c->sps.getQpBDOffset();
    c->sps.setChromaQpMappingTableFromParams( p );
    c->sps.deriveChromaQPMappingTables();
  }
  c->pps.setQpOffset( COMPONENT_Cb, s->cbQpOffset );
  c->pps.setQpOffset( COMPONENT_Cr, s->crQpOffset );
  c->pps.setQpOffset( JOINT_CbCr,   s->jointQpOffset );
  c->ph->setJointCbCrSignFlag( s->jointCbCrSign );
  c->slice.setDepQuantEnabledFlag( s->depQuant );
  c->slice.setExplicitScalingListUsed( false );

  const UnitArea picArea( fmt, Area( 0, 0, 128, 128 ) );
  c->cs.m_reco.create( picArea );
  c->cs.picHeader = c->ph;
  c->cs.area = picArea;

  CodingUnit& cu = c->cu;
  memset( (void*) &cu, 0, sizeof( cu ) );
  const UnitArea ua( fmt, Area( 0, 0, s->w, s->h ) );
  cu.UnitArea::operator=( ua );
  cu.firstTU.UnitArea::operator=( ua );
  cu.lastTU = &cu.firstTU;
  cu.cs = &c->cs; cu.slice = &c->slice; cu.pps = &c->pps; cu.sps = &c->sps;
  cu.qp = s->qp; cu.chromaQpAdj = s->chromaQpAdj;
  cu.setPredMode( s->predMode ? MODE_INTRA : MODE_INTER );
  cu.setLfnstIdx( s->lfnstIdx );
  cu.intraDir[0] = s->intraDirL; cu.intraDir[1] = s->intraDirC;
  cu.setMipFlag( s->mipFlag );
  cu.setIspMode( s->ispMode );
  cu.setSbtInfo( ( s->sbtIdx & 0xf ) | ( s->sbtPos << 4 ) );
  cu.setBdpcmMode( s->bdpcmL ); cu.setBdpcmModeChroma( s->bdpcmC );
  cu.setTreeType( s->sepTree ? TREE_C : TREE_D );
  cu.setChType( s->sepTree ? CH_C : CH_L );

  TransformUnit& tu = cu.firstTU;
  tu.cu = &cu; tu.next = nullptr;
  tu.setChType( cu.chType() );
  const ComponentID comp = ComponentID( s->comp );
  tu.jointCbCr = isChroma( comp ) ? s->jointCbCr : 0;
  tu.cbf = tu.jointCbCr ? ( tu.jointCbCr << 1 ) : ( 1 << s->comp );    // cbf bit per component: Y=1,Cb=2,Cr=4
  tu.setMtsIdx( s->comp, s->mtsIdx );
  tu.maxScanPosX[s->comp] = s->maxScanPosX; tu.maxScanPosY[s->comp] = s->maxScanPosY;

  const ComponentID coded = tu.jointCbCr ? b200glue::jccrCodedComp( tu ) : comp;
  if( tu.jointCbCr ) { tu.setMtsIdx( coded, s->mtsIdx ); tu.maxScanPosX[coded] = s->maxScanPosX; tu.maxScanPosY[coded] = s->maxScanPosY; }
  const CompArea& area = tu.blocks[coded];
  PelBuf plane = c->cs.getRecoBuf( area );
  for( unsigned y = 0; y < area.height; y++ ) for( unsigned x = 0; x < area.width; x++ ) plane.at( x, y ) = levels[y * area.width + x];

  // (a) our flattener
  std::vector<int16_t> coefs;
  TrQuant& tq = *sharedTQ();
  const bool made = b200glue::flattenTU( tu, coded, tq, coefs, *rec );
  *numCoefs = (int32_t) coefs.size();
  memcpy( coefsOut, coefs.data(), coefs.size() * sizeof( int16_t ) );

  // (b) the reference, as DecCu::reconstructResi (DecCu.cpp:536-577) drives it
  if( tu.jointCbCr )
  {
    PelBuf resiCb = c->cs.getRecoBuf( tu.blocks[COMPONENT_Cb] );
    PelBuf resiCr = c->cs.getRecoBuf( tu.blocks[COMPONENT_Cr] );
    if( tu.jointCbCr >> 1 ) { QpParam qp( tu, COMPONENT_Cb ); tq.invTransformNxN( tu, COMPONENT_Cb, resiCb, qp ); }
    else                    { QpParam qp( tu, COMPONENT_Cr ); tq.invTransformNxN( tu, COMPONENT_Cr, resiCr, qp ); }
    tq.invTransformICT( tu, resiCb, resiCr );
    PelBuf r0 = coded == COMPONENT_Cb ? resiCb : resiCr, r1 = coded == COMPONENT_Cb ? resiCr : resiCb;
    for( unsigned y = 0; y < area.height; y++ ) for( unsigned x = 0; x < area.width; x++ ) { resi0[y * area.width + x] = r0.at( x, y ); resi1[y * area.width + x] = r1.at( x, y ); }
  }
  else
  {
    QpParam qp( tu, comp );
    tq.invTransformNxN( tu, comp, plane, qp );
    for( unsigned y = 0; y < area.height; y++ ) for( unsigned x = 0; x < area.width; x++ ) resi0[y * area.width + x] = plane.at( x, y );
  }
  return made ? 1 : 0;
}

// ================================================================================================ fake picture
// A real vvdec Picture + CodingStructure, created through the reference's own calls (Picture::create,
// Picture::finalInit, allocateNewSlice), so that picture-level reference functions can run on synthetic state.
struct FakePicture
{
  std::shared_ptr<SPS> sps = std::make_shared<SPS>();
  std::shared_ptr<PPS> pps = std::make_shared<PPS>();
  std::shared_ptr<PicHeader> ph = std::make_shared<PicHeader>();
  CUChunkCache cuCache;
  TUChunkCache tuCache;
  Picture pic;
  std::vector<LoopFilterParam> lfp;
  int W, H, ctu, W4, H4;

  FakePicture( const b200_geom& g, int numSlices )
  {
    globalInit();
    W = g.width; H = g.height; ctu = g.ctuSize; W4 = ( W + 3 ) >> 2; H4 = ( H + 3 ) >> 2;
    const ChromaFormat fmt = g.chromaFormat ? CHROMA_420 : CHROMA_400;
    sps->setChromaFormatIdc( fmt );
    sps->setBitDepth( g.bitDepth );
    sps->setQpBDOffset( 6 * ( g.bitDepth - 8 ) );
    sps->setMaxPicWidthInLumaSamples( W ); sps->setMaxPicHeightInLumaSamples( H );
    sps->setCTUSize( ctu ); sps->setMaxCUWidth( ctu ); sps->setMaxCUHeight( ctu );
    sps->setLog2MinCodingBlockSize( 2 );
    pps->setPicWidthInLumaSamples( W ); pps->setPicHeightInLumaSamples( H );
    pps->pcv = std::make_unique<PreCalcValues>( *sps, *pps );
    pic.create( fmt, Size( W, H ), ctu, ctu + 16, 0 );
    const APS* noAps[ALF_CTB_MAX_NUM_APS] = { nullptr };
    pic.finalInit( &cuCache, &tuCache, sps.get(), pps.get(), ph, noAps, nullptr, nullptr, false );
    for( int i = 0; i < numSlices; i++ )
    {
      Slice* sl = pic.allocateNewSlice( nullptr );
      sl->setSPS( sps.get() ); sl->setPPS( pps.get() ); sl->setPicHeader( ph.get() );
      sl->setPic( &pic );
      sl->getClpRngs().bd = g.bitDepth;   // Slice.cpp:407
    }
    CodingStructure& cs = *pic.cs;
    const PreCalcValues& pcv = *cs.pcv;
    lfp.assign( (size_t) pcv.num4x4CtuBlks * pcv.sizeInCtus * 2, LoopFilterParam{} );
    for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
    {
      CtuData& cd = cs.getCtuData( a );
      cd.slice = pic.slices[0]; cd.pps = pps.get(); cd.sps = sps.get(); cd.ph = ph.get();
      cd.ctuIdx = a; cd.colIdx = a % pcv.widthInCtus; cd.lineIdx = a / pcv.widthInCtus;
      cd.lfParam[0] = lfp.data() + (size_t) a * pcv.num4x4CtuBlks;
      cd.lfParam[1] = lfp.data() + (size_t) ( pcv.sizeInCtus + a ) * pcv.num4x4CtuBlks;
    }
  }
  ~FakePicture() { pic.parseDone.unlock(); }

  void setPlanes( const b200_geom& g, int16_t* const planes[3] )
  {
    for( int c = 0; c < ( g.chromaFormat ? 3 : 1 ); c++ )
    {
      PelBuf b = pic.cs->getRecoBuf( ComponentID( c ) );
      for( unsigned y = 0; y < b.height; y++ ) memcpy( b.buf + y * b.stride, planes[c] + (size_t) y * g.stride[c], b.width * sizeof( Pel ) );
    }
  }
  void getPlanes( const b200_geom& g, int16_t* const planes[3], bool fromFlt = false, PelUnitBuf* flt = nullptr )
  {
    for( int c = 0; c < ( g.chromaFormat ? 3 : 1 ); c++ )
    {
      PelBuf b = flt ? flt->bufs[c] : pic.cs->getRecoBuf( ComponentID( c ) );
      for( unsigned y = 0; y < b.height; y++ ) memcpy( planes[c] + (size_t) y * g.stride[c], b.buf + y * b.stride, b.width * sizeof( Pel ) );
    }
  }
  // raster [H4][W4] grid -> the per-CTU arrays (stride = ctu/4) the reference walks (CodingStructure.h:229)
  void setLfGrid( int dir, const b200_lf_param* grid )
  {
    CodingStructure& cs = *pic.cs;
    const int c4 = ctu >> 2;
    for( int y4 = 0; y4 < H4; y4++ ) for( int x4 = 0; x4 < W4; x4++ )
    {
      const int a = cs.ctuRsAddr( x4 / c4, y4 / c4 );
      LoopFilterParam& d = cs.getCtuData( a ).lfParam[dir][( y4 % c4 ) * c4 + ( x4 % c4 )];
      static_assert( sizeof( LoopFilterParam ) == sizeof( b200_lf_param ), "layout" );
      memcpy( &d, &grid[y4 * W4 + x4], sizeof( d ) );
    }
  }
};

// ------------------------------------------------------------------------------------------------ K3
extern "C" void ref_lf_pel_filter_luma( int simd, int16_t* src, ptrdiff_t step, ptrdiff_t offset, int tc, int sw, int thrCut,
                                        int fsP, int fsQ, int bitDepth )
{
  static LoopFilter lfs( false ), lfv( true );
  ClpRng clp; clp.bd = bitDepth;
  ( simd ? lfv : lfs ).xPelFilterLuma( src, step, offset, tc, sw != 0, thrCut, fsP != 0, fsQ != 0, clp );
}

extern "C" void ref_lf_filtering_pq( int simd, int16_t* src, ptrdiff_t step, ptrdiff_t offset, int numP, int numQ, int tc )
{
  static LoopFilter lfs( false ), lfv( true );
  ( simd ? lfv : lfs ).xFilteringPandQ( src, step, offset, numP, numQ, tc );
}

extern "C" int ref_lf_deblock_picture( int simd, const b200_geom* g, int16_t* const planes[3], const b200_lf_param* lfV,
                                       const b200_lf_param* lfH, const uint8_t* ctuSlice, const b200_lf_slice* slices, int numSlices,
                                       const b200_lf_seq* seq, int dirs )
{
  FakePicture fp( *g, numSlices );
  CodingStructure& cs = *fp.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  if( seq && seq->ladfEnabled )
  {
    fp.sps->setLadfEnabled( true ); fp.sps->setLadfNumIntervals( seq->ladfNumIntervals );
    for( int k = 0; k < seq->ladfNumIntervals; k++ ) { fp.sps->setLadfQpOffset( seq->ladfQpOffset[k], k ); fp.sps->setLadfIntervalLowerBound( seq->ladfIntervalLowerBound[k], k ); }
  }
  for( int i = 0; i < numSlices; i++ )
  {
    Slice* sl = fp.pic.slices[i];
    sl->setDeblockingFilterDisable( slices[i].disable );
    sl->setDeblockingFilterBetaOffsetDiv2( slices[i].betaOffsetDiv2[0] ); sl->setDeblockingFilterTcOffsetDiv2( slices[i].tcOffsetDiv2[0] );
    sl->setDeblockingFilterCbBetaOffsetDiv2( slices[i].betaOffsetDiv2[1] ); sl->setDeblockingFilterCbTcOffsetDiv2( slices[i].tcOffsetDiv2[1] );
    sl->setDeblockingFilterCrBetaOffsetDiv2( slices[i].betaOffsetDiv2[2] ); sl->setDeblockingFilterCrTcOffsetDiv2( slices[i].tcOffsetDiv2[2] );
  }
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) cs.getCtuData( a ).slice = fp.pic.slices[ctuSlice ? ctuSlice[a] : 0];
  fp.setPlanes( *g, planes );
  fp.setLfGrid( 0, lfV ); fp.setLfGrid( 1, lfH );
  LoopFilter lf( simd != 0 );
  if( dirs & 1 ) for( unsigned y = 0; y < pcv.heightInCtus; y++ ) for( unsigned x = 0; x < pcv.widthInCtus; x++ ) lf.loopFilterCTU( cs, MAX_NUM_CHANNEL_TYPE, x, y, EDGE_VER );
  if( dirs & 2 ) for( unsigned y = 0; y < pcv.heightInCtus; y++ ) for( unsigned x = 0; x < pcv.widthInCtus; x++ ) lf.loopFilterCTU( cs, MAX_NUM_CHANNEL_TYPE, x, y, EDGE_HOR );
  fp.getPlanes( *g, planes );
  return 0;
}
