// ref_shim.cpp — OUR extern "C" access layer over the unmodified reference (linked from
// oracle/_ref/libvvdec_ref.a).  TEST INFRASTRUCTURE ONLY: pins oracle/*.c and the flattener
// (vvdec_b200/vvdec_glue) against the reference's own functions, and provides the "reference"
// CPU baseline for bench.py.  Same construction trick as tests/vvdec_unit_test/vvdec_unit_test.cpp.
#include "ref_shim.h"
#include "ref_stream.h"
#include <mutex>
#include <random>
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
#include "CommonLib/Reshape.h"
#include "CommonLib/SEI_internal.h"
#include "../vvdec_b200/vvdec_glue/flatten_pu.h"
#include "../vvdec_b200/vvdec_glue/flatten_filters.h"
#include "../vvdec_b200/vvdec_glue/DecLibReconB200.h"   // compile check of the drop-in class against the reference headers
#include <assert.h>
#include <fstream>
#include <chrono>
#include <ctime>
#include <iomanip>
#include "vvdec/vvdec.h"
#include "MD5.h"
namespace b200_ref_app {      // the application's static helpers (plane writers); every header they include is already in
#include "vvdecHelper.h"
}
#include "CommonLib/AdaptiveLoopFilter.h"
#include "CommonLib/RdCost.h"
#include <chrono>
#include <thread>
#include <atomic>
#include "../vvdec_b200/vvdec_glue/flatten_tu.h"
#include "CommonLib/IntraPrediction.h"
#include "../vvdec_b200/vvdec_glue/flatten_intra.h"
#include <memory>
#include "vvdec/sei.h"
#include "FilmGrain/FilmGrainImpl.h"
#define class struct            // FilmGrain's state is private by default class access: open it for the glue flattener (test code only)
#include "FilmGrain/FilmGrain.h"
#undef class
#include "../vvdec_b200/vvdec_glue/flatten_output.h"

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

extern "C" void ref_dequant_scaling( int simd, int width, int maxX, int maxY, int scaleQP, const int32_t* dequantCoef, const int16_t* q, size_t qStride, int32_t* coef,
                                     int rightShift, int inputMaximum, int32_t transformMaximum )
{
  globalInit();
  static Quant qs( nullptr, false ), qv( nullptr, true );
  ( simd ? qv : qs ).DeQuantScaling( width, maxX, maxY, scaleQP, dequantCoef, q, qStride, coef, rightShift, inputMaximum, transformMaximum );
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

// ------------------------------------------------------------------------------------------------ K4
extern "C" void ref_sao_offset_block( int simd, int bitDepth, int typeIdx, const int* offset32, int startIdx, const int16_t* src, int16_t* dst,
                                      ptrdiff_t srcStride, ptrdiff_t dstStride, int width, int height, unsigned avail,
                                      int numVerVb, const int* verVb, int numHorVb, const int* horVb )
{
  static SampleAdaptiveOffset ss( false ), sv( true );
  std::vector<int8_t> b1( width + 2 ), b2( width + 2 );
  int offs[MAX_NUM_SAO_CLASSES]; memcpy( offs, offset32, sizeof( offs ) );
  int hv[3] = { -1, -1, -1 }, vv[3] = { -1, -1, -1 };
  for( int i = 0; i < numVerVb; i++ ) vv[i] = verVb[i];
  for( int i = 0; i < numHorVb; i++ ) hv[i] = horVb[i];
  ClpRng clp; clp.bd = bitDepth;
  ( simd ? sv : ss ).offsetBlock( bitDepth, clp, typeIdx, offs, startIdx, src, dst, srcStride, dstStride, width, height,
                                   avail & B200_AVAIL_L, avail & B200_AVAIL_R, avail & B200_AVAIL_A, avail & B200_AVAIL_B,
                                   avail & B200_AVAIL_AL, avail & B200_AVAIL_AR, avail & B200_AVAIL_BL, avail & B200_AVAIL_BR,
                                   &b1, &b2, numVerVb + numHorVb > 0, hv, vv, numHorVb, numVerVb );
}

// one CU per CTU so that the picture-level filters find cuPtr[0][0] / slice / tile of every CTU
static void addCtuCUs( FakePicture& fp )
{
  CodingStructure& cs = *fp.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  for( unsigned y = 0; y < pcv.heightInCtus; y++ ) for( unsigned x = 0; x < pcv.widthInCtus; x++ )
  {
    const int w = std::min<int>( pcv.maxCUWidth, pcv.lumaWidth - x * pcv.maxCUWidth ), h = std::min<int>( pcv.maxCUHeight, pcv.lumaHeight - y * pcv.maxCUHeight );
    CodingUnit& cu = cs.addCU( UnitArea( pcv.chrFormat, Area( x * pcv.maxCUWidth, y * pcv.maxCUHeight, w, h ) ), CH_L, TREE_D, MODE_TYPE_ALL, nullptr, nullptr );
    cu.slice = fp.pic.slices[0]; cu.pps = fp.pps.get(); cu.sps = fp.sps.get(); cu.tileIdx = 0;
  }
}

extern "C" int ref_sao_picture( int simd, const b200_geom* g, const int16_t* const src[3], int16_t* const dst[3], const b200_sao_ctu* ctus, const b200_vb* vb )
{
  FakePicture fp( *g, 1 );
  CodingStructure& cs = *fp.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  fp.pps->setLoopFilterAcrossSlicesEnabledFlag( true ); fp.pps->setLoopFilterAcrossTilesEnabledFlag( true );
  if( vb && ( vb->numVer || vb->numHor ) )
  {
    fp.ph->setVirtualBoundariesPresentFlag( true );
    fp.ph->setNumVerVirtualBoundaries( vb->numVer ); fp.ph->setNumHorVirtualBoundaries( vb->numHor );
    for( int i = 0; i < vb->numVer; i++ ) fp.ph->setVirtualBoundariesPosX( vb->posX[i], i );
    for( int i = 0; i < vb->numHor; i++ ) fp.ph->setVirtualBoundariesPosY( vb->posY[i], i );
  }
  addCtuCUs( fp );
  int16_t* s3[3] = { (int16_t*) src[0], (int16_t*) src[1], (int16_t*) src[2] };
  fp.setPlanes( *g, s3 );
  PelStorage tmp; tmp.create( pcv.chrFormat, Size( g->width, g->height ), g->ctuSize, 16, MEMORY_ALIGN_DEF_SIZE );
  tmp.copyFrom( cs.getRecoBuf() );
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
  {
    SAOBlkParam& bp = cs.getCtuData( a ).saoParam;
    bp.reset();
    for( int c = 0; c < ( g->chromaFormat ? 3 : 1 ); c++ )
    {
      if( ctus[a].type[c] == B200_SAO_OFF ) continue;
      bp[c].modeIdc = SAO_MODE_NEW; bp[c].typeIdc = ctus[a].type[c]; bp[c].typeAuxInfo = ctus[a].band[c];
      if( ctus[a].type[c] == B200_SAO_BO ) for( int i = 0; i < 4; i++ ) bp[c].offset[( ctus[a].band[c] + i ) & 31] = ctus[a].offset[c][i];
      else for( int i = 0; i < 5; i++ ) bp[c].offset[i] = ctus[a].offset[c][i];
    }
  }
  SampleAdaptiveOffset sao( simd != 0 );
  sao.create( g->width, g->height, pcv.chrFormat, g->ctuSize, g->ctuSize, 0, 0, tmp );
  for( unsigned y = 0; y < pcv.heightInCtus; y++ ) for( unsigned x = 0; x < pcv.widthInCtus; x++ )
    sao.SAOProcessCTU( cs, clipArea( UnitArea( pcv.chrFormat, Area( x * g->ctuSize, y * g->ctuSize, g->ctuSize, g->ctuSize ) ), fp.pic ) );
  fp.getPlanes( *g, dst );
  return 0;
}

// ------------------------------------------------------------------------------------------------ K5
static PelStorage wrapPlane( const int16_t* p, ptrdiff_t stride, int w, int h, int pad )
{
  // padded private copy (the reference's filters read up to 3/4 samples beyond the block; callers extend borders)
  PelStorage st; st.create( CHROMA_400, Size( w, h ), 0, pad, MEMORY_ALIGN_DEF_SIZE );
  PelBuf b = st.bufs[0];
  for( int y = 0; y < h; y++ ) memcpy( b.buf + y * b.stride, p + y * stride, w * sizeof( Pel ) );
  st.extendBorderPel( pad );
  return st;
}

extern "C" void ref_alf_classify( int simd, uint16_t* cls, const int16_t* srcLuma, ptrdiff_t stride, int planeW, int planeH,
                                  int blkX, int blkY, int blkW, int blkH, int shift, int vbCtuHeight, int vbPos )
{
  static AdaptiveLoopFilter as( false ), av( true );
  PelStorage st = wrapPlane( srcLuma, stride, planeW, planeH, 8 );
  AlfClassifier c[64];
  memset( c, 0, sizeof( c ) );
  ( simd ? av : as ).m_deriveClassificationBlk( c, st.bufs[0], Area( blkX, blkY, blkW, blkH ), shift, vbCtuHeight, vbPos );
  for( int i = 0; i < 64; i++ ) cls[i] = c[i].classIdx | ( c[i].transposeIdx << 8 );
}

extern "C" void ref_alf_filter_blk( int simd, int is7x7, const uint16_t* cls, int16_t* dst, ptrdiff_t dstStride, const int16_t* src, ptrdiff_t srcStride,
                                    int planeW, int planeH, int blkX, int blkY, int blkW, int blkH, const int16_t* coeff, const int16_t* clip,
                                    int bitDepth, int vbCtuHeight, int vbPos )
{
  static AdaptiveLoopFilter as( false ), av( true );
  AdaptiveLoopFilter& a = simd ? av : as;
  // the pointers take Unit buffers and a component id: put the plane into a 4:0:0 (luma) or 4:2:0 (chroma as Cb) unit
  const bool chroma = !is7x7;
  PelStorage s, d;
  if( chroma ) { s.create( CHROMA_420, Size( planeW * 2, planeH * 2 ), 0, 16, MEMORY_ALIGN_DEF_SIZE ); d.create( CHROMA_420, Size( planeW * 2, planeH * 2 ), 0, 16, MEMORY_ALIGN_DEF_SIZE ); }
  else         { s.create( CHROMA_400, Size( planeW, planeH ), 0, 8, MEMORY_ALIGN_DEF_SIZE );          d.create( CHROMA_400, Size( planeW, planeH ), 0, 8, MEMORY_ALIGN_DEF_SIZE ); }
  const ComponentID comp = chroma ? COMPONENT_Cb : COMPONENT_Y;
  PelBuf sb = s.bufs[comp], db = d.bufs[comp];
  for( int y = 0; y < planeH; y++ ) { memcpy( sb.buf + y * sb.stride, src + y * srcStride, planeW * sizeof( Pel ) ); memcpy( db.buf + y * db.stride, dst + y * dstStride, planeW * sizeof( Pel ) ); }
  s.extendBorderPel( chroma ? 8 : 8 );
  AlfClassifier c[64];
  if( cls ) for( int i = 0; i < 64; i++ ) c[i] = AlfClassifier( cls[i] & 0xff, cls[i] >> 8 );
  ClpRng clp; clp.bd = bitDepth;
  ( is7x7 ? a.m_filter7x7Blk : a.m_filter5x5Blk )( cls ? c : nullptr, d, s, Area( blkX, blkY, blkW, blkH ), comp, coeff, clip, clp, vbCtuHeight, vbPos, false );
  for( int y = 0; y < planeH; y++ ) memcpy( dst + y * dstStride, db.buf + y * db.stride, planeW * sizeof( Pel ) );
}

extern "C" void ref_alf_ccalf_blk( int simd, int16_t* dstChroma, ptrdiff_t chromaStride, const int16_t* srcLuma, ptrdiff_t lumaStride,
                                   int lumaW, int lumaH, int cX, int cY, int cW, int cH, const int16_t* coeff, int bitDepth, int vbCtuHeight, int vbPos )
{
  static AdaptiveLoopFilter as( false ), av( true );
  PelStorage s; s.create( CHROMA_420, Size( lumaW, lumaH ), 0, 16, MEMORY_ALIGN_DEF_SIZE );
  PelBuf lb = s.bufs[0];
  for( int y = 0; y < lumaH; y++ ) memcpy( lb.buf + y * lb.stride, srcLuma + y * lumaStride, lumaW * sizeof( Pel ) );
  s.extendBorderPel( 8 );
  PelBuf dbuf( dstChroma, chromaStride, lumaW >> 1, lumaH >> 1 );
  ClpRng clp; clp.bd = bitDepth;
  ( simd ? av : as ).m_filterCcAlf( dbuf, s, Area( cX, cY, cW, cH ), Area( cX * 2, cY * 2, cW * 2, cH * 2 ), COMPONENT_Cb, coeff, clp, vbCtuHeight, vbPos );
}

// APS / slice / CTU state of the ALF stage from the flattened tables (the inverse of b200glue::buildAlfTables / flattenALF)
static void setupAlfSlice( FakePicture& fp, Slice* sl, const b200_alf_ctu* ctus, const b200_alf_tables* T )
{
  CodingStructure& cs = *fp.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  // APS 0..n-1 carry the luma sets 16.. ; APS 7 carries the chroma alternatives (<=8) ; APS 6/5 the CC-ALF filters (<=4 each)
  static std::shared_ptr<APS> apsStore[ALF_CTB_MAX_NUM_APS];
  const APS* apss[ALF_CTB_MAX_NUM_APS] = { nullptr };
  for( int i = 0; i < ALF_CTB_MAX_NUM_APS; i++ ) { apsStore[i] = std::make_shared<APS>(); apsStore[i]->setAPSId( i ); apss[i] = apsStore[i].get(); }
  const int nApsLuma = T->numLumaSets - NUM_FIXED_FILTER_SETS;
  AlfApsIdVec ids;
  for( int i = 0; i < nApsLuma; i++ )
  {
    AlfSliceParam& p = apsStore[i]->getAlfAPSParam();
    memcpy( p.lumaCoeffFinal, T->lumaCoeff + (size_t) ( 16 + i ) * 1300, 1300 * sizeof( short ) );
    memcpy( p.lumaClippFinal, T->lumaClip  + (size_t) ( 16 + i ) * 1300, 1300 * sizeof( short ) );
    p.lumaFinalDone = true;
    ids.push_back( i );
  }
  sl->setNumAlfAps( nApsLuma ); sl->setAlfApsIdsLuma( ids );
  {
    AlfSliceParam& p = apsStore[7]->getAlfAPSParam();
    p.numAlternativesChroma = T->numChromaAlts;
    memcpy( p.chromaCoeff,    T->chromaCoeff, T->numChromaAlts * 7 * sizeof( short ) );
    memcpy( p.chrmClippFinal, T->chromaClip,  T->numChromaAlts * 7 * sizeof( short ) );
    p.chrmFinalDone = true;
    sl->setAlfApsIdChroma( 7 );
  }
  for( int c = 0; c < 2; c++ )
  {
    CcAlfFilterParam& p = apsStore[6 - c]->getCcAlfAPSParam();
    for( int k = 0; k < T->numCc[c]; k++ ) memcpy( p.ccAlfCoeff[c][k], T->ccCoeff[c] + k * 7, 7 * sizeof( short ) );
    p.ccAlfFilterCount[c] = (uint8_t) T->numCc[c];
  }
  sl->setCcAlfCbEnabledFlag( T->numCc[0] > 0 ); sl->setCcAlfCrEnabledFlag( T->numCc[1] > 0 );
  sl->setCcAlfCbApsId( 6 ); sl->setCcAlfCrApsId( 5 );
  sl->setAlfApss( apss );
  for( int c = 0; c < 3; c++ ) sl->setAlfEnabledFlag( ComponentID( c ), true );
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
  {
    CtuAlfData& d = cs.getCtuData( a ).alfParam;
    for( int c = 0; c < 3; c++ ) d.alfCtuEnableFlag[c] = ctus[a].enable[c] & 1;
    d.alfCtbFilterIndex = ctus[a].lumaSet;
    for( int c = 0; c < 2; c++ ) { d.alfCtuAlternative[c] = ctus[a].chromaAlt[c]; d.ccAlfFilterControl[c] = ctus[a].ccIdx[c]; }
  }
}

extern "C" int ref_alf_picture( int simd, const b200_geom* g, const int16_t* const src[3], int16_t* const dst[3], const b200_alf_ctu* ctus,
                                const b200_alf_tables* T )
{
  FakePicture fp( *g, 1 );
  CodingStructure& cs = *fp.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  fp.pps->setLoopFilterAcrossSlicesEnabledFlag( true ); fp.pps->setLoopFilterAcrossTilesEnabledFlag( true );
  fp.sps->setUseALF( true ); fp.sps->setUseCCALF( true );
  addCtuCUs( fp );
  Slice* sl = fp.pic.slices[0];
  { SliceMap sm; sm.addCtusToSlice( 0, pcv.widthInCtus, 0, pcv.heightInCtus, pcv.widthInCtus ); sl->setSliceMap( sm ); }
  setupAlfSlice( fp, sl, ctus, T );
  int16_t* s3[3] = { (int16_t*) src[0], (int16_t*) src[1], (int16_t*) src[2] };
  fp.setPlanes( *g, s3 );
  PelStorage out; out.create( pcv.chrFormat, Size( g->width, g->height ), g->ctuSize, 16, MEMORY_ALIGN_DEF_SIZE );
  AdaptiveLoopFilter alf( simd != 0 );
  alf.create( fp.ph.get(), fp.sps.get(), fp.pps.get(), 1, out );
  for( unsigned y = 0; y < pcv.heightInCtus; y++ ) for( unsigned x = 0; x < pcv.widthInCtus; x++ ) alf.prepareCTU( cs, x, y );
  for( unsigned y = 0; y < pcv.heightInCtus; y++ ) for( unsigned x = 0; x < pcv.widthInCtus; x++ ) alf.processCTU( cs, x, y, 0 );
  fp.getPlanes( *g, dst, true, &out );
  return 0;
}

// ------------------------------------------------------------------------------------------------ K2
// GEO CU from a record: partition p uses DPB slot refSlot[p] = list (slot >> 1), refIdx (slot & 1) in the shims' reference-list layout
static void setupGeo( CodingUnit& cu, const b200_pu& pu )
{
  cu.setGeoFlag( true ); cu.geoSplitDir = (uint8_t) pu.bcwW1; cu.setMergeFlag( true ); cu.setBcwIdx( BCW_DEFAULT ); cu.setSmvdMode( 0 ); cu.setAffineFlag( false );
  cu.setInterDirrefIdxGeo0( uint8_t( ( ( ( pu.refSlot[0] >> 1 ) + 1 ) << 4 ) | ( pu.refSlot[0] & 1 ) ) );
  cu.setInterDirrefIdxGeo1( uint8_t( ( ( ( pu.refSlot[1] >> 1 ) + 1 ) << 4 ) | ( pu.refSlot[1] & 1 ) ) );
  cu.mv[0][1] = Mv( pu.mv[0][0], pu.mv[0][1] ); cu.mv[1][1] = Mv( pu.mv[1][0], pu.mv[1][1] );
}
static void runMc( InterPrediction& ip, CodingUnit& cu, PelUnitBuf& buf )
{
  if( cu.geoFlag() ) ip.motionCompensationGeo( cu, buf ); else ip.motionCompensation( cu, buf, true, true );
}

static const int32_t* g_refWpRaw = nullptr;
extern "C" void ref_set_wp( const int32_t* raw ) { g_refWpRaw = raw; }
static void applyWp( FakePicture& cur, Slice* sl )
{
  cur.pps->setWPBiPred( g_refWpRaw != nullptr );
  if( !g_refWpRaw ) return;
  for( int l = 0; l < 2; l++ ) for( int i = 0; i < 2; i++ ) for( int c = 0; c < 3; c++ )
  {
    const int32_t* r = g_refWpRaw + ( ( l * 2 + i ) * 3 + c ) * 3;
    WPScalingParam& w = sl->m_weightPredTable[l][i][c];
    w.uiLog2WeightDenom = r[0]; w.iWeight = r[1]; w.iOffset = r[2];
    w.bPresentFlag = r[1] != ( 1 << r[0] ) || r[2] != 0;
  }
}

extern "C" int ref_mc_predict( int simd, const b200_geom* g, int16_t* const dst[3], const int16_t* const* refs, const b200_pu* pus, size_t numPus,
                               int32_t* dmvrMv, size_t numDmvr )
{
  FakePicture cur( *g, 1 );
  std::unique_ptr<FakePicture> ref[4];
  for( int s = 0; s < 4; s++ )
  {
    ref[s].reset( new FakePicture( *g, 1 ) );
    int16_t* p3[3] = { (int16_t*) refs[s * 3], (int16_t*) refs[s * 3 + 1], (int16_t*) refs[s * 3 + 2] };
    ref[s]->setPlanes( *g, p3 );
    ref[s]->pic.extendPicBorder();
  }
  CodingStructure& cs = *cur.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  SPS& sps = *cur.sps;
  sps.setUseBIO( true ); sps.setUseDMVR( true ); sps.setUseBcw( true ); sps.setUseAffine( true ); sps.setUseAffineType( true ); sps.setUsePROF( true );
  Slice* sl = cur.pic.slices[0];
  sl->setSliceType( B_SLICE ); sl->setPOC( 8 );
  const int pocs[4] = { 4, 0, 12, 16 };
  for( int l = 0; l < 2; l++ ) for( int i = 0; i < 2; i++ )
  {
    sl->m_apcRefPicList[l][i] = &ref[l * 2 + i]->pic;
    sl->m_aiRefPOCList [l][i] = pocs[l * 2 + i];
    sl->m_bIsUsedAsLongTerm[l][i] = false;
    ref[l * 2 + i]->pic.poc = pocs[l * 2 + i];
  }
  sl->setNumRefIdx( REF_PIC_LIST_0, 2 ); sl->setNumRefIdx( REF_PIC_LIST_1, 2 );
  sl->resetWpScaling();
  applyWp( cur, sl );

  std::vector<MotionInfo> motion( (size_t) pcv.num4x4CtuBlks * pcv.sizeInCtus );
  std::vector<Mv> dmvrCache( (size_t) pcv.num8x8CtuBlks * pcv.sizeInCtus + 16 );
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) cs.getCtuData( a ).motion = motion.data() + (size_t) a * pcv.num4x4CtuBlks;
  cs.m_dmvrMvCache = dmvrCache.data();

  static RdCost rdScalar( false ), rdSimd( true );
  std::unique_ptr<InterPrediction> ip( new InterPrediction() );
  ip->init( simd ? &rdSimd : &rdScalar, pcv.chrFormat, g->ctuSize, simd != 0 );

  PelStorage predStore; predStore.create( pcv.chrFormat, Size( 128, 128 ), 0, 0, MEMORY_ALIGN_DEF_SIZE );
  int rc = 0;
  for( size_t n = 0; n < numPus; n++ )
  {
    const b200_pu& pu = pus[n];
    const UnitArea ua( pcv.chrFormat, Area( pu.x, pu.y, pu.w, pu.h ) );
    CodingUnit& cu = cs.addCU( ua, CH_L, TREE_D, MODE_TYPE_ALL, nullptr, nullptr );
    cu.slice = sl; cu.pps = cur.pps.get(); cu.sps = cur.sps.get();
    cu.setPredMode( MODE_INTER );
    for( int l = 0; l < 2; l++ )
    {
      cu.refIdx[l] = pu.refSlot[l] < 0 ? -1 : ( pu.refSlot[l] & 1 );
      cu.mv[l][0] = Mv( pu.mv[l][0], pu.mv[l][1] );
      cu.mv[l][1] = Mv( pu.cpmv[l][0][0], pu.cpmv[l][0][1] );
      cu.mv[l][2] = Mv( pu.cpmv[l][1][0], pu.cpmv[l][1][1] );
    }
    cu.setInterDir( pu.interDir );
    cu.setImv( ( pu.flags & B200_PU_ALTHPEL ) ? IMV_HPEL : IMV_OFF );
    int bcwIdx = BCW_DEFAULT;
    for( int i = 0; i < BCW_NUM; i++ ) if( getBcwWeight( g_BcwInternBcw[i], REF_PIC_LIST_1 ) == pu.bcwW1 ) bcwIdx = i;
    cu.setBcwIdx( bcwIdx );
    const bool wantDmvr = pu.flags & B200_PU_DMVR, wantBio = pu.flags & B200_PU_BDOF, affine = pu.flags & B200_PU_AFFINE;
    cu.setMergeFlag( wantDmvr );          // DMVR needs a regular merge CU; without it bio alone is decided by POC distances / size
    cu.setMergeType( MRG_TYPE_DEFAULT_N );
    cu.setSmvdMode( ( !wantBio && !affine && pu.refSlot[0] >= 0 && pu.refSlot[1] >= 0 && pu.bcwW1 == 4 ) ? 1 : 0 );   // smvd switches BDOF off without touching the pixels
    cu.setAffineFlag( affine );
    cu.setAffineType( ( pu.flags & B200_PU_AFFINE6 ) ? AFFINEMODEL_6PARAM : AFFINEMODEL_4PARAM );
    cur.ph->setDisProfFlag( affine && !( pu.flags & ( B200_PU_PROF0 | B200_PU_PROF1 ) ) );
    if( affine ) { for( int l = 0; l < 2; l++ ) if( cu.refIdx[l] >= 0 ) PU::setAllAffineMv( cu, cu.mv[l][0], cu.mv[l][1], cu.mv[l][2], RefPicList( l ) ); }
    if( pu.flags & B200_PU_GEO ) setupGeo( cu, pu ); else PU::spanMotionInfo( cu );

    PelUnitBuf predBuf = predStore.getBuf( UnitArea( pcv.chrFormat, Area( 0, 0, pu.w, pu.h ) ) );
    runMc( *ip, cu, predBuf );
    if( (bool) cu.dmvrCondition() != wantDmvr ) rc = -1;
    for( int c = 0; c < ( g->chromaFormat ? 3 : 1 ); c++ )
    {
      const PelBuf& b = predBuf.bufs[c];
      for( unsigned y = 0; y < b.height; y++ ) memcpy( dst[c] + (size_t) ( ( pu.y >> ( c ? 1 : 0 ) ) + y ) * g->stride[c] + ( pu.x >> ( c ? 1 : 0 ) ), b.buf + y * b.stride, b.width * sizeof( Pel ) );
    }
    if( wantDmvr && dmvrMv )
    {
      const int nSub = std::max( 1, pu.w >> 4 ) * std::max( 1, pu.h >> 4 );
      for( int k = 0; k < nSub && pu.dmvrOff + k < numDmvr; k++ ) { const Mv& m = cs.m_dmvrMvCache[cu.mvdL0SubPuOff + k]; dmvrMv[( pu.dmvrOff + k ) * 2] = m.hor; dmvrMv[( pu.dmvrOff + k ) * 2 + 1] = m.ver; }
    }
  }
  return rc;
}

// ================================================================================================ filter flatteners (vvdec_glue/flatten_filters.h)
// Fills the reference structures from flattened input (as the filter shims above do), then runs the flatteners on them: what comes
// back must be the input.  SAO availability comes from the real deriveLoopFilterBoundaryAvailibility.
extern "C" int ref_flatten_filters( const b200_geom* g, const b200_lf_param* lfV, const b200_lf_param* lfH, const b200_sao_ctu* sao, const b200_alf_ctu* alf,
                                    const b200_alf_tables* T, b200_lf_param* lfVOut, b200_lf_param* lfHOut, b200_sao_ctu* saoOut, b200_alf_ctu* alfOut,
                                    int16_t* lumaCoeffOut, int16_t* lumaClipOut, int16_t* chromaCoeffOut, int16_t* chromaClipOut, int16_t* cc0Out, int16_t* cc1Out, int32_t counts[4] )
{
  FakePicture fp( *g, 1 );
  CodingStructure& cs = *fp.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  fp.pps->setLoopFilterAcrossSlicesEnabledFlag( true ); fp.pps->setLoopFilterAcrossTilesEnabledFlag( true );
  fp.sps->setUseALF( true ); fp.sps->setUseCCALF( true );
  addCtuCUs( fp );
  Slice* sl = fp.pic.slices[0];
  { SliceMap sm; sm.addCtusToSlice( 0, pcv.widthInCtus, 0, pcv.heightInCtus, pcv.widthInCtus ); sl->setSliceMap( sm ); }
  fp.setLfGrid( 0, lfV ); fp.setLfGrid( 1, lfH );
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) { b200glue::flattenLfCtu( cs, a, 0, lfVOut ); b200glue::flattenLfCtu( cs, a, 1, lfHOut ); }
  PelStorage flt; flt.create( pcv.chrFormat, Size( g->width, g->height ), g->ctuSize, 16, MEMORY_ALIGN_DEF_SIZE );
  SampleAdaptiveOffset saoF( false );
  saoF.create( g->width, g->height, pcv.chrFormat, g->ctuSize, g->ctuSize, 0, 0, flt );
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
  {
    SAOBlkParam& bp = cs.getCtuData( a ).saoParam; bp.reset();
    for( int c = 0; c < 3; c++ )
    {
      if( sao[a].type[c] == B200_SAO_OFF ) continue;
      bp[c].modeIdc = SAO_MODE_NEW; bp[c].typeIdc = sao[a].type[c]; bp[c].typeAuxInfo = sao[a].band[c];
      if( sao[a].type[c] == B200_SAO_BO ) for( int i = 0; i < 4; i++ ) bp[c].offset[( sao[a].band[c] + i ) & 31] = sao[a].offset[c][i];
      else for( int i = 0; i < 5; i++ ) bp[c].offset[i] = sao[a].offset[c][i];
    }
    bool av[8];
    saoF.deriveLoopFilterBoundaryAvailibility( cs, Position( ( a % pcv.widthInCtus ) * g->ctuSize, ( a / pcv.widthInCtus ) * g->ctuSize ), av[0], av[1], av[2], av[3], av[4], av[5], av[6], av[7] );
    b200glue::flattenSAO( bp, av, g->chromaFormat ? 3 : 1, saoOut[a] );
  }
  setupAlfSlice( fp, sl, alf, T );
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) b200glue::flattenALF( cs.getCtuData( a ).alfParam, alfOut[a] );
  AdaptiveLoopFilter alfF( false );
  alfF.create( fp.ph.get(), fp.sps.get(), fp.pps.get(), 1, flt );      // fills m_clipDefault for the bit depth
  b200glue::AlfTableStore store;
  const b200_alf_tables R = b200glue::buildAlfTables( *sl, &alfF.m_fixedFilterSetCoeffDec[0][0], alfF.m_clipDefault, store );
  memcpy( lumaCoeffOut, R.lumaCoeff, store.lumaCoeff.size() * 2 ); memcpy( lumaClipOut, R.lumaClip, store.lumaClip.size() * 2 );
  memcpy( chromaCoeffOut, R.chromaCoeff, store.chromaCoeff.size() * 2 ); memcpy( chromaClipOut, R.chromaClip, store.chromaClip.size() * 2 );
  memcpy( cc0Out, R.ccCoeff[0], store.cc[0].size() * 2 ); memcpy( cc1Out, R.ccCoeff[1], store.cc[1].size() * 2 );
  counts[0] = R.numLumaSets; counts[1] = R.numChromaAlts; counts[2] = R.numCc[0]; counts[3] = R.numCc[1];
  return 0;
}

// ================================================================================================ output writer of vvdecapp
extern "C" size_t ref_write_component( const int16_t* src, ptrdiff_t stride, int w, int h, int fmt, uint8_t* dst, size_t cap )
{
  vvdecPlane pl; memset( &pl, 0, sizeof( pl ) );
  // vvdecPlane::stride is in bytes (vvdec.h:459) and the pyuv loop halves it (:128); the 8-bit loop adds it to an unsigned short* (:93), i.e.
  // walks it as samples — feed each branch the unit its pointer arithmetic uses
  pl.ptr = (unsigned char*) src; pl.width = w; pl.height = h; pl.stride = (uint32_t) ( fmt == 2 ? stride : stride * 2 ); pl.bytesPerSample = 2;
  std::ostringstream os;
  b200_ref_app::_writeComponentToFile( &os, &pl, nullptr, fmt == 2 ? 1 : 2, fmt == 1 );
  const std::string out = os.str();
  memcpy( dst, out.data(), std::min( cap, out.size() ) );
  return out.size();
}

// defined (non-static, but not declared in any header) in CommonLib/PicYuvMD5.cpp:138,:179,:198
namespace vvdec {
uint32_t calcCRC( const CPelUnitBuf& pic, PictureHash& digest, const BitDepths& bitDepths );
uint32_t calcChecksum( const CPelUnitBuf& pic, PictureHash& digest, const BitDepths& bitDepths );
uint32_t calcMD5( const CPelUnitBuf& pic, PictureHash& digest, const BitDepths& bitDepths );
}
extern "C" int ref_picture_hash( int method, int bitDepth, int16_t* const planes[3], const ptrdiff_t strides[3], int w, int h, uint8_t* digest, int cap )
{
  PelUnitBuf ub;
  ub.chromaFormat = CHROMA_420;
  for( int c = 0; c < 3; c++ ) ub.bufs.push_back( PelBuf( planes[c], strides[c], c ? w >> 1 : w, c ? h >> 1 : h ) );
  BitDepths bd; bd.recon = bitDepth;
  PictureHash dg;
  if( method == 1 ) calcCRC( ub, dg, bd ); else if( method == 2 ) calcChecksum( ub, dg, bd ); else calcMD5( ub, dg, bd );
  const int n = (int) dg.hash.size();
  for( int i = 0; i < n && i < cap; i++ ) digest[i] = dg.hash[i];
  return n;
}

// ================================================================================================ intra prediction
// A picture whose CUs (single tree, one TU each, decoding order) are all intra; the planes hold the reconstruction of the neighbourhood.
// The LAST CU of the list is predicted with the real IntraPrediction, exactly as DecCu::predAndReco drives it (DecCu.cpp:329-371), into the
// planes; its records come back through the glue flattener, and the flattener's availability is checked against m_neighborSize.
static int intraPredictCu( IntraPrediction& ip, CodingStructure& cs, CodingUnit& cu, const b200_geom* g, const int16_t* const resi[3], bool addResi, b200_intra_tu* recs, int capRecs, int& n )
{
  TransformUnit& tu = cu.firstTU;
  if( cu.ciipFlag() )
  {
    // the planes hold the CU's inter prediction; predBlendIntraCiip (DecCu.cpp:450-453) blends the planar intra prediction into it
    for( int c = 0; c < 3; c++ )
    {
      b200_intra_tu r;
      if( b200glue::flattenCiipBlock( cu, ComponentID( c ), r ) != b200glue::FLATTEN_INTRA_OK ) continue;
      if( addResi && resi && resi[c] ) r.flags |= B200_INTRA_ADD_RESI;
      if( n < capRecs ) recs[n] = r;
      n++;
    }
    PelUnitBuf predUnit = cs.getRecoBuf( cu );
    ip.predBlendIntraCiip( predUnit, cu );
    if( addResi && resi )
      for( const CompArea& area : cu.blocks )
      {
        if( !area.valid() || ( !isLuma( area.compID() ) && cu.chromaSize().width <= 2 ) ) continue;
        PelBuf p = cs.getRecoBuf( area ); const int pmax = ( 1 << g->bitDepth ) - 1, c = area.compID();
        for( unsigned y = 0; y < area.height; y++ ) for( unsigned x = 0; x < area.width; x++ )
          p.buf[y * p.stride + x] = (Pel) std::min( pmax, std::max( 0, p.buf[y * p.stride + x] + resi[c][( area.y + y ) * g->stride[c] + area.x + x] ) );
      }
    return 0;
  }
  for( const CompArea& area : tu.blocks )
  {
    if( !area.valid() ) continue;
    const ComponentID compID = area.compID();
    b200_intra_tu r;
    if( b200glue::flattenIntraTU( tu, compID, r ) != b200glue::FLATTEN_INTRA_OK ) return -2;
    PelBuf piPred = cs.getRecoBuf( area );
    const bool filt = isLuma( compID ) && cu.ispMode() == NOT_INTRA_SUBPARTITIONS && IntraPrediction::useFilteredIntraRefSamples( compID, cu, tu );
    ip.initIntraPatternChType( tu, area, filt );
    if( ( ( r.flags & B200_INTRA_AVAIL_TL ) ? 1 : 0 ) != ip.m_neighborSize[0] || r.numAbove != ip.m_neighborSize[1] || r.numLeft != ip.m_neighborSize[2] )
    {
      fprintf( stderr, "ref_intra: availability mismatch comp %d: glue (%d %d %d) reference (%d %d %d)\n", (int) compID, ( r.flags >> 1 ) & 1, r.numAbove, r.numLeft,
               ip.m_neighborSize[0], ip.m_neighborSize[1], ip.m_neighborSize[2] );
      return -3;
    }
    if( CU::isMIP( cu, toChannelType( compID ) ) )
    {                                                          // DecCu.cpp:321-326
      ip.initIntraPatternChType( tu, area );
      ip.initIntraMip( cu, area );
      ip.predIntraMip( compID, piPred, cu );
    }
    else if( compID != COMPONENT_Y && PU::isLMCMode( PU::getFinalIntraMode( cu, toChannelType( compID ) ) ) )
    {                                                          // DecCu.cpp:327-332
      ip.initIntraPatternChType( tu, area );
      ip.xGetLumaRecPixels( cu, area );
      ip.predIntraChromaLM( compID, piPred, cu, area, PU::getFinalIntraMode( cu, toChannelType( compID ) ) );
    }
    else ip.predIntraAng( compID, piPred, cu, filt );
    if( addResi && resi && resi[compID] )
    {                                                          // piReco.reconstruct( piPred, piResi, clpRng ) (DecCu.cpp:392)
      r.flags |= B200_INTRA_ADD_RESI;
      const int pmax = ( 1 << g->bitDepth ) - 1;
      for( unsigned y = 0; y < area.height; y++ ) for( unsigned x = 0; x < area.width; x++ )
      {
        Pel& p = piPred.buf[y * piPred.stride + x];
        p = (Pel) std::min( pmax, std::max( 0, p + resi[compID][( area.y + y ) * g->stride[compID] + area.x + x] ) );
      }
    }
    if( n < capRecs ) recs[n] = r;
    n++;
  }
  return 0;
}

// ISP luma of one CU, as DecCu::predAndReco walks it (DecCu.cpp:284-398): per sub-TU initIntraPatternChTypeISP (or once per 4-wide prediction
// region), predIntraAng, pred + residual.  rec returns what the oracle needs: CU geometry, mode, the CU-level availability counts, and in
// lmLeft / lmAbove whether the CU has a left / above neighbour.
static int intraPredictIspCu( IntraPrediction& ip, CodingStructure& cs, CodingUnit& cu, const b200_geom* g, const int16_t* const resi[3], unsigned resiMask, b200_intra_tu& rec )
{
  memset( &rec, 0, sizeof( rec ) );
  int k = 0;
  for( TransformUnit& tu : TUTraverser( &cu.firstTU, cu.lastTU->next ) )
  {
    const CompArea& area = tu.blocks[COMPONENT_Y];
    const bool predRegDiffFromTB = CU::isPredRegDiffFromTB( cu, COMPONENT_Y );
    const bool firstTBInPredReg  = CU::isFirstTBInPredReg( cu, COMPONENT_Y, area );
    CompArea areaPredReg( COMPONENT_Y, area );
    PelBuf piReco = cs.getRecoBuf( area );
    if( predRegDiffFromTB ) { if( firstTBInPredReg ) { CU::adjustPredArea( areaPredReg ); ip.initIntraPatternChTypeISP( cu, areaPredReg, piReco ); } }
    else ip.initIntraPatternChTypeISP( cu, area, piReco );
    if( k == 0 )
    {
      rec.x = (uint16_t) cu.lx(); rec.y = (uint16_t) cu.ly(); rec.log2w = (uint8_t) getLog2( cu.lwidth() ); rec.log2h = (uint8_t) getLog2( cu.lheight() );
      rec.mode = (uint8_t) PU::getFinalIntraMode( cu, CH_L ); rec.flags = ip.m_neighborSize[0] ? B200_INTRA_AVAIL_TL : 0;
      rec.numAbove = (uint8_t) ip.m_neighborSize[1]; rec.numLeft = (uint8_t) ip.m_neighborSize[2];
      rec.lmLeft  = nullptr != cs.getCURestricted( cu.lumaPos().offset( -1, 0 ), cu, CH_L, cu.left );
      rec.lmAbove = nullptr != cs.getCURestricted( cu.lumaPos().offset( 0, -1 ), cu, CH_L, cu.above );
    }
    if( !predRegDiffFromTB || firstTBInPredReg )
    {
      PelBuf piPred = cs.getRecoBuf( areaPredReg );
      ip.predIntraAng( COMPONENT_Y, piPred, cu, false );
    }
    if( resi && resi[0] && ( ( resiMask >> k ) & 1 ) )
    {
      const int pmax = ( 1 << g->bitDepth ) - 1;
      for( unsigned y = 0; y < area.height; y++ ) for( unsigned x = 0; x < area.width; x++ )
      {
        Pel& p = piReco.buf[y * piReco.stride + x];
        p = (Pel) std::min( pmax, std::max( 0, p + resi[0][( area.y + y ) * g->stride[0] + area.x + x] ) );
      }
    }
    k++;
  }
  return k;
}

// all != 0: every CU of the list is predicted (and, with rsv[1] set and resi given, reconstructed) in order — a whole intra picture
extern "C" int ref_intra_case( int simd, const b200_geom* g, int16_t* const planes[3], const int16_t* const resi[3], const ref_intra_cu* cus, int numCus, int all,
                               b200_intra_tu* recs, int capRecs, int cclmCollocated )
{
  try
  {
    FakePicture cur( *g, 1 );
    cur.setPlanes( *g, planes );
    CodingStructure& cs = *cur.pic.cs;
    const PreCalcValues& pcv = *cs.pcv;
    Slice* sl = cur.pic.slices[0];
    sl->setSliceType( I_SLICE );
    cur.sps->setVerCollocatedChromaFlag( cclmCollocated != 0 );        // = getCclmCollocatedChromaFlag (Slice.h:1793)
    IntraPrediction ip;
    ip.init( pcv.chrFormat, g->bitDepth );                              // x86: installs the SIMD kernels (IntraPrediction.cpp:399)
    if( !simd )
    {
      IntraPrediction scalar;                                           // the constructor installs the C kernels (:362-376)
      ip.IntraPredAngleCore4 = scalar.IntraPredAngleCore4; ip.IntraPredAngleCore8 = scalar.IntraPredAngleCore8;
      ip.IntraPredAngleChroma4 = scalar.IntraPredAngleChroma4; ip.IntraPredAngleChroma8 = scalar.IntraPredAngleChroma8;
      ip.IntraPredSampleFilter8 = scalar.IntraPredSampleFilter8; ip.IntraPredSampleFilter16 = scalar.IntraPredSampleFilter16;
      ip.xPredIntraPlanar = scalar.xPredIntraPlanar; ip.GetLumaRecPixel420 = scalar.GetLumaRecPixel420;
    }
    int n = 0;
    for( int i = 0; i < numCus; i++ )
    {
      const ref_intra_cu& c = cus[i];
      UnitArea ua( pcv.chrFormat, Area( c.x, c.y, c.w, c.h ) );
      if( c.rsv[0] ) { ua.blocks[1].width = ua.blocks[1].height = 0; ua.blocks[2].width = ua.blocks[2].height = 0; }   // luma CU of a local dual tree (4xN, 8x4)
      const Position pos( c.x, c.y );
      const CodingUnit* left  = cs.getCURestricted( pos.offset( -1, 0 ), pos, 0, 0, CH_L );       // as CABACReader::coding_tree does for every CU
      const CodingUnit* above = cs.getCURestricted( pos.offset( 0, -1 ), pos, 0, 0, CH_L );
      CodingUnit& cu = cs.addCU( ua, CH_L, c.rsv[0] ? TREE_L : TREE_D, c.rsv[0] ? MODE_TYPE_INTRA : MODE_TYPE_ALL, left, above );
      cu.slice = sl; cu.pps = cur.pps.get(); cu.sps = cur.sps.get();
      cu.setPredMode( ( c.rsv[2] & 12 ) ? MODE_INTER : MODE_INTRA );                                    // bit 2: CIIP CU, bit 3: plain inter CU (a neighbour)
      cu.intraDir[0] = c.dirL; cu.intraDir[1] = c.dirC;
      if( c.rsv[2] & 4 ) { cu.setCiipFlag( true ); cu.intraDir[0] = PLANAR_IDX; cu.intraDir[1] = DM_CHROMA_IDX; }
      cu.setMultiRefIdx( c.multiRefIdx ); cu.setBdpcmMode( c.bdpcm ); cu.setBdpcmModeChroma( c.bdpcmC );
      if( c.rsv[2] & 1 ) { cu.setMipFlag( true ); cu.setMipTransposedFlag( ( c.rsv[2] & 2 ) != 0 ); }      // dirL is the MIP mode index then
      const int isp = ( c.rsv[2] >> 4 ) & 3;                                                                // 1 horizontal, 2 vertical split into sub-partitions
      if( isp )
      {
        // sub-TUs like PartitionerImpl::getTUIntraSubPartitions (UnitPartitioner.cpp:628): luma strips, the chroma blocks ride in the last one
        cu.setIspMode( isp );
        const int part = (int) CU::getISPSplitDim( c.w, c.h, isp == 1 ? TU_1D_HORZ_SPLIT : TU_1D_VERT_SPLIT ), nParts = ( isp == 1 ? c.h : c.w ) / part;
        for( int k = 0; k < nParts; k++ )
        {
          UnitArea sub = ua;
          CompArea& y = sub.blocks[COMPONENT_Y];
          if( isp == 1 ) { y.height = part; y.y = c.y + k * part; } else { y.width = part; y.x = c.x + k * part; }
          if( k + 1 < nParts ) { sub.blocks[1] = CompArea(); sub.blocks[2] = CompArea(); }
          cs.addTU( sub, CH_L, cu );
        }
        if( all || i == numCus - 1 )
        {
          b200_intra_tu r;
          if( intraPredictIspCu( ip, cs, cu, g, resi, c.bdpcmC /* residual mask of the partitions */, r ) <= 0 ) return -4;
          if( n < capRecs ) recs[n] = r;
          n++;
        }
        continue;
      }
      cs.addTU( ua, CH_L, cu );
      if( ( all || i == numCus - 1 ) && !( c.rsv[2] & 8 ) )                                                // plain inter CUs: their samples are given
        if( int rc = intraPredictCu( ip, cs, cu, g, resi, c.rsv[1] != 0, recs, capRecs, n ) ) return rc;
    }
    cur.getPlanes( *g, planes );
    return n;
  }
  catch( std::exception& e ) { fprintf( stderr, "ref_intra_case: %s\n", e.what() ); return -9; }
}

// ================================================================================================ film grain
// The real FilmGrain (firmware + SIMD line kernels) over `frames` consecutive frames, as VVDecImpl::xAddGrain drives it (vvdecimpl.cpp:898);
// the tables and line seeds of the LAST frame come back through the glue flattener, so that oracle / device can redo that frame.
// scalarImpl selects FilmGrainImpl (the C model) instead of the SIMD class the decoder runs on x86: they differ where a scale LUT entry is
// >= 128 at 10 bit (FilmGrainImpl_X86_SIMD.h:450,:478 sign-extend the uint8 scale; the C model :282 keeps it unsigned).
// sei: modelId, log2ScaleFactor, then per component: present, numModelValues, numIntervals, then per interval: lower, upper, 6 model values.
static std::unique_ptr<FilmGrain> g_fg;
static b200glue::FilmGrainTables  g_fgTabs;
extern "C" int ref_film_grain( const int* sei, int scalarImpl, int bitDepth, int w, int h, int frames, int16_t* const planes[3], const ptrdiff_t strides[3],
                               int8_t* pattern, uint8_t* sLUT, uint8_t* pLUT, uint32_t* lineSeeds, int* scaleShift, uint8_t* compPresent )
{
  try
  {
    auto fgc = std::make_unique<vvdecSEIFilmGrainCharacteristics>();
    memset( fgc.get(), 0, sizeof( *fgc ) );
    fgc->filmGrainModelId = (uint8_t) *sei++; fgc->log2ScaleFactor = (uint8_t) *sei++;
    for( int c = 0; c < 3; c++ )
    {
      vvdecCompModel& cm = fgc->compModel[c];
      cm.presentFlag = *sei++ != 0; cm.numModelValues = (uint8_t) *sei++; cm.numIntensityIntervals = (uint16_t) *sei++;
      for( int i = 0; i < cm.numIntensityIntervals; i++ )
      {
        cm.intensityValues[i].intensityIntervalLowerBound = (uint8_t) *sei++; cm.intensityValues[i].intensityIntervalUpperBound = (uint8_t) *sei++;
        for( int v = 0; v < 6; v++ ) cm.intensityValues[i].compModelValue[v] = *sei++;
      }
    }
    g_fg = std::make_unique<FilmGrain>();                    // constructs the SIMD implementation on x86 (FilmGrain.cpp:722)
    if( scalarImpl ) g_fg->m_impl = std::make_unique<FilmGrainImpl>();   // the C model (FilmGrainImpl.cpp)
    g_fg->updateFGC( fgc.get() );
    const int bytes = bitDepth > 8 ? 2 : 1;
    std::vector<uint8_t> buf[3];
    for( int f = 0; f < frames; f++ )
    {
      g_fg->setDepth( bitDepth );
      g_fg->setColorFormat( planes[1] ? VVDEC_CF_YUV420_PLANAR : VVDEC_CF_YUV400_PLANAR );
      g_fg->prepareBlockSeeds( w, h );
      if( f + 1 < frames ) continue;                     // earlier frames only advance the seed state
      g_fgTabs.flatten( *g_fg );
      // frame copy with the sample size of the output frame (8-bit frames are byte planes, xCreateFrame vvdecimpl.cpp:1238) and a row pitch with
      // slack: the line kernels always finish the last 16-sample block, also when it sticks out of the picture
      uint8_t* base[3] = { nullptr, nullptr, nullptr }; ptrdiff_t sb[3] = { 0, 0, 0 };
      const int nPl = planes[1] ? 3 : 1;
      for( int c = 0; c < nPl; c++ )
      {
        const int cw = c ? w >> 1 : w, ch = c ? h >> 1 : h;
        sb[c] = (ptrdiff_t) ( cw + 64 ) * bytes; buf[c].assign( (size_t) sb[c] * ( ch + 1 ), 0 ); base[c] = buf[c].data();
        for( int y = 0; y < ch; y++ ) for( int x = 0; x < cw; x++ )
        {
          if( bytes == 2 ) ( (uint16_t*) ( base[c] + y * sb[c] ) )[x] = (uint16_t) planes[c][y * strides[c] + x];
          else             base[c][y * sb[c] + x] = (uint8_t) planes[c][y * strides[c] + x];
        }
      }
      for( int y = 0; y < h; y++ )
        g_fg->add_grain_line( base[0] + sb[0] * y, base[1] ? base[1] + sb[1] * ( y / 2 ) : nullptr, base[2] ? base[2] + sb[2] * ( y / 2 ) : nullptr, y, w );
      for( int c = 0; c < nPl; c++ )
      {
        const int cw = c ? w >> 1 : w, ch = c ? h >> 1 : h;
        for( int y = 0; y < ch; y++ ) for( int x = 0; x < cw; x++ )
          planes[c][y * strides[c] + x] = bytes == 2 ? (int16_t) ( (uint16_t*) ( base[c] + y * sb[c] ) )[x] : (int16_t) base[c][y * sb[c] + x];
      }
    }
    memcpy( pattern, g_fgTabs.pattern.data(), g_fgTabs.pattern.size() );
    memcpy( sLUT, g_fgTabs.sLUT, 768 ); memcpy( pLUT, g_fgTabs.pLUT, 768 );
    memcpy( lineSeeds, g_fgTabs.lineSeeds.data(), g_fgTabs.lineSeeds.size() * 4 );
    *scaleShift = g_fgTabs.fg.scaleShift;
    for( int c = 0; c < 3; c++ ) compPresent[c] = g_fgTabs.fg.compPresent[c];
    return 0;
  }
  catch( std::exception& e ) { fprintf( stderr, "ref_film_grain: %s\n", e.what() ); return -1; }
}

// ================================================================================================ LMCS (Reshape)
static std::unique_ptr<Reshape> g_rsp;
static PelBufferOps g_pelOpsScalar, g_pelOpsSimd; static bool g_pelOpsInit = false;
static void initPelOps()
{
  if( g_pelOpsInit ) return;
  g_pelOpsScalar = PelBufferOps();
  g_pelOpsSimd = PelBufferOps();
#if ENABLE_SIMD_OPT_BUFFER && defined( TARGET_SIMD_X86 )
  g_pelOpsSimd.initPelBufOpsX86();
#endif
  g_pelOpsInit = true;
}

static void fillReshapeFromTables( Reshape& r, int bitDepth, const b200_lmcs* L )
{
  r.createDec( bitDepth );
  r.m_sliceReshapeInfo.sliceReshaperEnableFlag = true; r.m_sliceReshapeInfo.sliceReshaperModelPresentFlag = true;
  r.m_sliceReshapeInfo.enableChromaAdj = L->chromaAdj;
  r.m_sliceReshapeInfo.reshaperModelMinBinIdx = L->minBinIdx; r.m_sliceReshapeInfo.reshaperModelMaxBinIdx = L->maxBinIdx;
  r.m_initCW = (uint16_t) L->orgCW;
  for( int i = 0; i < 17; i++ ) { r.m_reshapePivot[i] = L->reshapePivot[i]; r.m_inputPivot[i] = L->inputPivot[i]; }
  for( int i = 0; i < 16; i++ )
  {
    r.m_fwdScaleCoef[i] = L->fwdScaleCoef[i]; r.m_chromaAdjHelpLUT[i] = L->chromaAdjHelpLUT[i];
    const int binCW = L->reshapePivot[i + 1] - L->reshapePivot[i];                       // constructReshaper :341-353
    r.m_binCW[i] = (uint16_t) binCW; r.m_invScaleCoef[i] = binCW ? Pel( L->orgCW * ( 1 << FP_PREC ) / binCW ) : Pel( 0 );
  }
  memcpy( r.m_invLUT, L->invLUT, sizeof( Pel ) << bitDepth );
}

extern "C" int ref_lmcs_build( int bitDepth, int minBin, int maxBin, const int* deltaCW, int chrResScalingOffset, int chromaAdj, b200_lmcs* out, int16_t* invLut )
{
  globalInit();
  g_rsp.reset( new Reshape() );
  Reshape& r = *g_rsp;
  r.createDec( bitDepth );
  SliceReshapeInfo& si = r.m_sliceReshapeInfo;
  si.sliceReshaperEnableFlag = true; si.sliceReshaperModelPresentFlag = true; si.enableChromaAdj = chromaAdj;
  si.reshaperModelMinBinIdx = minBin; si.reshaperModelMaxBinIdx = maxBin; si.chrResScalingOffset = chrResScalingOffset;
  for( int i = 0; i < PIC_CODE_CW_BINS; i++ ) si.reshaperModelBinCWDelta[i] = deltaCW[i];
  try { r.constructReshaper(); } catch( ... ) { return -1; }
  memset( out, 0, sizeof( *out ) );
  out->chromaAdj = chromaAdj; out->minBinIdx = minBin; out->maxBinIdx = maxBin; out->orgCW = r.m_initCW;
  for( int i = 0; i < 17; i++ ) { out->reshapePivot[i] = r.m_reshapePivot[i]; out->inputPivot[i] = r.m_inputPivot[i]; }
  for( int i = 0; i < 16; i++ ) { out->fwdScaleCoef[i] = r.m_fwdScaleCoef[i]; out->chromaAdjHelpLUT[i] = r.m_chromaAdjHelpLUT[i]; }
  memcpy( invLut, r.m_invLUT, sizeof( Pel ) << bitDepth );
  out->invLUT = invLut;
  return 0;
}

extern "C" void ref_lmcs_fwd_block( int simd, int16_t* ptr, ptrdiff_t stride, int w, int h )
{
  initPelOps();
  const PelBufferOps saved = g_pelBufOP; g_pelBufOP = simd ? g_pelOpsSimd : g_pelOpsScalar;
  PelBuf b( ptr, stride, w, h );
  g_rsp->rspBufFwd( b );
  g_pelBufOP = saved;
}

extern "C" void ref_lmcs_inv_block( int simd, int16_t* ptr, ptrdiff_t stride, int w, int h )
{
  initPelOps();
  const Reshape& r = *g_rsp;
  const PelBufferOps& ops = simd ? g_pelOpsSimd : g_pelOpsScalar;
  // the two branches of Reshape::rspCtuBcw (:399-406)
  if( ops.rspBcw ) ops.rspBcw( ptr, stride, w, h, r.m_lumaBD, r.m_sliceReshapeInfo.reshaperModelMinBinIdx, r.m_sliceReshapeInfo.reshaperModelMaxBinIdx, r.m_reshapePivot.data(), r.m_invScaleCoef.data(), r.m_inputPivot.data() );
  else             ops.applyLut( ptr, stride, w, h, r.m_invLUT );
}

extern "C" void ref_lmcs_scale_block( int16_t* ptr, ptrdiff_t stride, int w, int h, int scale, int bitDepth )
{
  PelBuf b( ptr, stride, w, h );
  ClpRng rng; rng.bd = bitDepth;
  b.scaleSignal( scale, rng );
}

extern "C" int ref_lmcs_vpdu_scale( const b200_geom* g, int16_t* const planes[3], int x, int y )
{
  FakePicture fp( *g, 1 );
  addCtuCUs( fp );
  fp.setPlanes( *g, planes );
  CodingStructure& cs = *fp.pic.cs;
  TransformUnit tu; memset( (void*) &tu, 0, sizeof( tu ) );
  tu.cu = cs.getCU( Position( x, y ), CH_L );
  g_rsp->setVPDULoc( -1, -1 );
  return g_rsp->calculateChromaAdjVpduNei( tu, Position( x, y ) );
}

// ================================================================================================ whole back end, multi-threaded
// Persistent worker pool (the reference keeps its worker threads alive across pictures too: Utilities/ThreadPool.h); items are handed out
// dynamically through an atomic counter.
struct WorkerPool
{
  std::vector<std::thread> workers;
  std::mutex m; std::condition_variable cvGo, cvDone;
  std::function<void( int, int )> job; int n = 0; std::atomic<int> next{ 0 }; int gen = 0, running = 0; bool stop = false;
  explicit WorkerPool( int threads )
  {
    for( int t = 0; t < threads; t++ ) workers.emplace_back( [this, t] {
      int seen = 0;
      for( ;; )
      {
        { std::unique_lock<std::mutex> lk( m ); cvGo.wait( lk, [&] { return stop || gen != seen; } ); if( stop ) return; seen = gen; }
        for( int i = next++; i < n; i = next++ ) job( i, t );
        { std::lock_guard<std::mutex> lk( m ); if( --running == 0 ) cvDone.notify_one(); }
      }
    } );
  }
  ~WorkerPool() { { std::lock_guard<std::mutex> lk( m ); stop = true; } cvGo.notify_all(); for( auto& w : workers ) w.join(); }
  template<class F> void run( int count, F f )
  {
    { std::lock_guard<std::mutex> lk( m ); job = f; n = count; next = 0; running = (int) workers.size(); gen++; }
    cvGo.notify_all();
    std::unique_lock<std::mutex> lk( m ); cvDone.wait( lk, [&] { return running == 0; } );
  }
};
template<class F> static void parallelFor( int n, WorkerPool& pool, F f ) { pool.run( n, f ); }

extern "C" int ref_flatten_pu_case( int simd, const b200_geom* g, const int16_t* const* refs, int altRefs, const ref_cu_syntax* cus, int numCus,
                                    int16_t* const dst[3], b200_pu* recs, int capRecs, int32_t* dmvrMv, int numDmvr )
{
  FakePicture cur( *g, 1 );
  std::unique_ptr<FakePicture> ref[4];
  for( int s = 0; s < 4; s++ )
  {
    ref[s].reset( new FakePicture( *g, 1 ) );
    int16_t* p3[3] = { (int16_t*) refs[s * 3], (int16_t*) refs[s * 3 + 1], (int16_t*) refs[s * 3 + 2] };
    ref[s]->setPlanes( *g, p3 );
    ref[s]->pic.extendPicBorder();
  }
  CodingStructure& cs = *cur.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  SPS& sps = *cur.sps;
  sps.setUseBIO( true ); sps.setUseDMVR( true ); sps.setUseBcw( true ); sps.setUseAffine( true ); sps.setUseAffineType( true ); sps.setUsePROF( true );
  Slice* sl = cur.pic.slices[0];
  sl->setSliceType( B_SLICE ); sl->setPOC( 8 );
  const int slotOf[2][2] = { { 0, 1 }, { 2, altRefs ? 0 : 3 } }, pocOf[4] = { 4, 0, 12, 16 };
  b200glue::SlotMap sm; memset( &sm, -1, sizeof( sm ) );
  for( int s = 0; s < 4; s++ ) ref[s]->pic.poc = pocOf[s];
  for( int l = 0; l < 2; l++ ) for( int i = 0; i < 2; i++ )
  {
    sl->m_apcRefPicList[l][i] = &ref[slotOf[l][i]]->pic; sl->m_aiRefPOCList[l][i] = pocOf[slotOf[l][i]]; sl->m_bIsUsedAsLongTerm[l][i] = false;
    sm.slot[l][i] = (int8_t) slotOf[l][i];
  }
  sl->setNumRefIdx( REF_PIC_LIST_0, 2 ); sl->setNumRefIdx( REF_PIC_LIST_1, 2 );
  sl->resetWpScaling();
  applyWp( cur, sl );
  std::vector<MotionInfo> motion( (size_t) pcv.num4x4CtuBlks * pcv.sizeInCtus );
  std::vector<Mv> dmvrCache( (size_t) std::max<size_t>( numDmvr + 64, (size_t) pcv.num8x8CtuBlks * pcv.sizeInCtus + 16 ) );
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) cs.getCtuData( a ).motion = motion.data() + (size_t) a * pcv.num4x4CtuBlks;
  cs.m_dmvrMvCache = dmvrCache.data();
  static RdCost rdScalar( false ), rdSimd( true );
  std::unique_ptr<InterPrediction> ip( new InterPrediction() );
  ip->init( simd ? &rdSimd : &rdScalar, pcv.chrFormat, g->ctuSize, simd != 0 );
  PelUnitBuf reco = cs.getRecoBuf();
  // (refIdx0, refIdx1) -> 1-based b200_wp index, the order synth.gen_wp uses
  auto wpIdxOf = []( int r0, int r1 ) { int k = 0; for( int a = -1; a <= 1; a++ ) for( int b = -1; b <= 1; b++ ) { if( a < 0 && b < 0 ) continue; k++; if( a == r0 && b == r1 ) return k; } return 0; };
  int n = 0; uint32_t dmvrOff = 0;
  for( int i = 0; i < numCus; i++ )
  {
    const ref_cu_syntax& c = cus[i];
    const UnitArea ua( pcv.chrFormat, Area( c.x, c.y, c.w, c.h ) );
    CodingUnit& cu = cs.addCU( ua, CH_L, TREE_D, MODE_TYPE_ALL, nullptr, nullptr );
    cu.slice = sl; cu.pps = cur.pps.get(); cu.sps = cur.sps.get();
    cu.setPredMode( MODE_INTER );
    for( int l = 0; l < 2; l++ ) { cu.refIdx[l] = c.refIdx[l]; for( int k = 0; k < 3; k++ ) cu.mv[l][k] = Mv( c.mv[l][k][0], c.mv[l][k][1] ); }
    cu.setInterDir( ( c.refIdx[0] >= 0 ? 1 : 0 ) + ( c.refIdx[1] >= 0 ? 2 : 0 ) );
    cu.setImv( c.imvHpel ? IMV_HPEL : IMV_OFF );
    cu.setBcwIdx( c.bcwIdx ); cu.setMergeFlag( c.mergeFlag ); cu.setMmvdFlag( c.mmvdFlag ); cu.setSmvdMode( c.smvd );
    cu.setMergeType( c.sbTmvp ? MRG_TYPE_SUBPU_ATMVP : MRG_TYPE_DEFAULT_N );
    cu.setAffineFlag( c.affine ); cu.setAffineType( c.affine6 ? AFFINEMODEL_6PARAM : AFFINEMODEL_4PARAM );
    cu.mvdL0SubPuOff = dmvrOff;
    if( c.w >= 8 && c.h >= 8 && c.w * c.h >= 128 ) dmvrOff += std::max( 1, c.w >> 4 ) * std::max( 1, c.h >> 4 );
    if( c.geo ) { cu.setGeoFlag( true ); cu.geoSplitDir = (uint8_t) c.geoSplitDir; cu.setInterDirrefIdxGeo0( (uint8_t) c.geoDir0 ); cu.setInterDirrefIdxGeo1( (uint8_t) c.geoDir1 ); cu.setMergeFlag( true ); }
    else if( c.affine ) { for( int l = 0; l < 2; l++ ) if( cu.refIdx[l] >= 0 ) PU::setAllAffineMv( cu, cu.mv[l][0], cu.mv[l][1], cu.mv[l][2], RefPicList( l ) ); }
    else PU::spanMotionInfo( cu );
    if( c.sbTmvp )
    {
      // seeded 8x8 motion field with few distinct candidates, so that runs of equal motion occur
      std::mt19937 rng( c.sbSeed );
      MotionInfo cand[3];
      for( auto& m : cand ) { const int dir = 1 + rng() % 3; for( int l = 0; l < 2; l++ ) { m.miRefIdx[l] = ( dir >> l ) & 1 ? int8_t( rng() % 2 ) : int8_t( MI_NOT_VALID ); m.mv[l] = Mv( int( rng() % 257 ) - 128, int( rng() % 257 ) - 128 ); } }
      for( int y = c.y; y < c.y + c.h; y += 8 ) for( int x = c.x; x < c.x + c.w; x += 8 )
      {
        const MotionInfo& m = cand[rng() % 3];
        for( int yy = 0; yy < 8 && y + yy < c.y + c.h; yy += 4 ) for( int xx = 0; xx < 8 && x + xx < c.x + c.w; xx += 4 )
          cs.getCtuData( cs.ctuRsAddr( Position( x + xx, y + yy ), CH_L ) ).motion[cs.inCtuPos( Position( x + xx, y + yy ), CH_L )] = m;
      }
    }
    PelUnitBuf predBuf = reco.subBuf( ua );
    runMc( *ip, cu, predBuf );
    b200glue::FlattenPuResult rc = b200glue::FLATTEN_PU_OK;
    if( c.sbTmvp ) rc = b200glue::flattenSbTmvp( cu, sm, wpIdxOf, [&]( const b200_pu& r ) { if( n < capRecs ) recs[n] = r; n++; } );
    else { b200_pu r; rc = b200glue::flattenPU( cu, sm, wpIdxOf, r ); if( rc == b200glue::FLATTEN_PU_OK ) { if( n < capRecs ) recs[n] = r; n++; } }
    if( rc != b200glue::FLATTEN_PU_OK ) return -1 - i;
  }
  cur.getPlanes( *g, dst );
  if( dmvrMv ) for( int i = 0; i < numDmvr; i++ ) { dmvrMv[2 * i] = dmvrCache[i].hor; dmvrMv[2 * i + 1] = dmvrCache[i].ver; }
  return n;
}

static void fillCuFromPu( CodingUnit& cu, const b200_pu& pu, FakePicture& cur, Slice* sl )
{
  memset( (void*) &cu, 0, sizeof( cu ) );
  const UnitArea ua( cur.pic.cs->pcv->chrFormat, Area( pu.x, pu.y, pu.w, pu.h ) );
  cu.UnitArea::operator=( ua );
  cu.cs = cur.pic.cs.get(); cu.slice = sl; cu.pps = cur.pps.get(); cu.sps = cur.sps.get();
  cu.ctuData = &cu.cs->getCtuData( cu.cs->ctuRsAddr( Position( pu.x, pu.y ), CH_L ) );
  cu.setChType( CH_L ); cu.setTreeType( TREE_D ); cu.setModeType( MODE_TYPE_ALL );
  cu.setPredMode( MODE_INTER );
  cu.mvdL0SubPuOff = pu.dmvrOff;
  for( int l = 0; l < 2; l++ )
  {
    cu.refIdx[l] = pu.refSlot[l] < 0 ? -1 : ( pu.refSlot[l] & 1 );
    cu.mv[l][0] = Mv( pu.mv[l][0], pu.mv[l][1] ); cu.mv[l][1] = Mv( pu.cpmv[l][0][0], pu.cpmv[l][0][1] ); cu.mv[l][2] = Mv( pu.cpmv[l][1][0], pu.cpmv[l][1][1] );
  }
  cu.setInterDir( pu.interDir );
  cu.setImv( ( pu.flags & B200_PU_ALTHPEL ) ? IMV_HPEL : IMV_OFF );
  int bcwIdx = BCW_DEFAULT;
  for( int i = 0; i < BCW_NUM; i++ ) if( getBcwWeight( g_BcwInternBcw[i], REF_PIC_LIST_1 ) == pu.bcwW1 ) bcwIdx = i;
  cu.setBcwIdx( bcwIdx );
  const bool wantDmvr = pu.flags & B200_PU_DMVR, wantBio = pu.flags & B200_PU_BDOF, affine = pu.flags & B200_PU_AFFINE;
  cu.setMergeFlag( wantDmvr ); cu.setMergeType( MRG_TYPE_DEFAULT_N );
  cu.setSmvdMode( ( !wantBio && !affine && pu.refSlot[0] >= 0 && pu.refSlot[1] >= 0 && pu.bcwW1 == 4 ) ? 1 : 0 );
  cu.setAffineFlag( affine );
  cu.setAffineType( ( pu.flags & B200_PU_AFFINE6 ) ? AFFINEMODEL_6PARAM : AFFINEMODEL_4PARAM );
  if( affine ) { for( int l = 0; l < 2; l++ ) if( cu.refIdx[l] >= 0 ) PU::setAllAffineMv( cu, cu.mv[l][0], cu.mv[l][1], cu.mv[l][2], RefPicList( l ) ); }
  if( pu.flags & B200_PU_GEO ) setupGeo( cu, pu ); else PU::spanMotionInfo( cu );
}

// K1 for one record with the reference's kernels (glue restated from TrQuant.cpp:201-485 / Quant.cpp:295-381)
static void refTuResidual( const b200_tu& tu, int bitDepth, const int16_t* coefs, TrQuant& tq, Quant& qnt, TCoeffOps& ops, TCoeff* dq, TCoeff* tmp, TCoeff* blk, Pel* resi, ptrdiff_t stride )
{
  const int w = 1 << tu.log2w, h = 1 << tu.log2h;
  int maxX = tu.maxX, maxY = tu.maxY;
  const int inputMaximum = ( 1 << ( tu.inBits - 1 ) ) - 1;
  const int16_t* q = coefs + tu.coefOff;
  memset( dq, 0, sizeof( TCoeff ) * w * h );
  if( tu.flags & ( B200_TU_BDPCM_H | B200_TU_BDPCM_V ) )
  {
    for( int y = 0; y < h; y++ ) for( int x = 0; x < w; x++ )
    {
      int v = q[y * ( maxX + 1 ) + x];
      if( ( tu.flags & B200_TU_BDPCM_H ) && x > 0 ) v = Clip3( -32768, 32767, dq[y * w + x - 1] + v );
      if( ( tu.flags & B200_TU_BDPCM_V ) && y > 0 ) v = Clip3( -32768, 32767, dq[( y - 1 ) * w + x] + v );
      dq[y * w + x] = v;
    }
    qnt.DeQuantPCM( w, w - 1, h - 1, tu.scale, dq, w, dq, tu.rightShift, inputMaximum, 32767 );
  }
  else
  {
    // the reference reads the levels in place from the reco plane (zero outside the coded corner, CABACReader.cpp:2457); its SIMD
    // DeQuant loads groups of 4/8 levels per row, so the packed corner is first laid out with the TU's stride like the plane has it
    alignas( 32 ) TCoeffSig lv[64 * 64];
    const int lw = std::max( w, 8 );
    memset( lv, 0, sizeof( TCoeffSig ) * lw * ( maxY + 1 ) );
    for( int y = 0; y <= maxY; y++ ) memcpy( lv + y * lw, q + y * ( maxX + 1 ), ( maxX + 1 ) * sizeof( TCoeffSig ) );
    qnt.DeQuant( w, maxX, maxY, tu.scale, lv, lw, dq, tu.rightShift, inputMaximum, 32767 );
  }
  if( tu.lfnst && !( tu.flags & B200_TU_TS ) )
  {
    static const uint8_t sx[16] = { 0,0,1,0,1,2,0,1,2,3,1,2,3,2,3,3 }, sy[16] = { 0,1,0,2,1,0,3,2,1,0,3,2,1,3,2,3 };
    const bool big = w >= 8 && h >= 8; const int sb = big ? 8 : 4, tr = ( tu.lfnst >> 4 ) & 1;
    int in[16], out[48];
    for( int i = 0; i < 16; i++ ) in[i] = dq[sy[i] * w + sx[i]];
    tq.m_invLfnstNxN( in, out, ( tu.lfnst >> 2 ) & 3, ( tu.lfnst & 3 ) - 1, sb, ( ( w == 4 && h == 4 ) || ( w == 8 && h == 8 ) ) ? 8 : 16 );
    const int* o = out;
    if( tr ) { if( sb == 4 ) { for( int y = 0; y < 4; y++ ) for( int x = 0; x < 4; x++ ) dq[y * w + x] = o[x * 4 + y]; }
               else for( int y = 0; y < 8; y++ ) { for( int x = 0; x < 4; x++ ) dq[y * w + x] = o[x * 8 + y]; if( y < 4 ) for( int x = 4; x < 8; x++ ) dq[y * w + x] = o[32 + ( x - 4 ) * 4 + y]; } }
    else for( int y = 0; y < sb; y++ ) { const int n = y < 4 ? sb : 4; for( int x = 0; x < n; x++ ) dq[y * w + x] = *o++; }
    maxX = std::max( maxX, std::min( w - 1, 7 ) ); maxY = std::max( maxY, std::min( h - 1, 7 ) );
  }
  if( tu.flags & B200_TU_TS ) { for( int y = 0; y < h; y++ ) for( int x = 0; x < w; x++ ) resi[y * stride + x] = Pel( dq[y * w + x] ); return; }
  const int trH = tu.trType & 3, trV = ( tu.trType >> 2 ) & 3, shift1 = 7, shift2 = 20 - bitDepth;
  if( maxX == 0 && maxY == 0 && trH == DCT2 && trV == DCT2 )
  {
    int dc = ( dq[0] * 64 + 64 ) >> shift1; dc = ( dc * 64 + ( 1 << ( shift2 - 1 ) ) ) >> shift2;
    for( int y = 0; y < h; y++ ) for( int x = 0; x < w; x++ ) resi[y * stride + x] = Pel( dc );
    return;
  }
  const int skipW = std::max( ( trH != DCT2 && w == 32 ) ? 16 : w > 32 ? w - 32 : 0, w - maxX - 1 );
  const int skipH = std::max( ( trV != DCT2 && h == 32 ) ? 16 : h > 32 ? h - 32 : 0, h - maxY - 1 );
  fastInvTrans[trV][tu.log2h - 1]( dq,  tmp, shift1, w, skipW, skipH, true,  -32768, 32767 );
  fastInvTrans[trH][tu.log2w - 1]( tmp, blk, shift2, h, 0,     skipW, false, -32768, 32767 );
  ops.cpyResiClip[tu.log2w]( blk, resi, stride, w, h, -32768, 32767, 1 << ( shift2 - 1 ), shift2 );
}

extern "C" double ref_decompress_picture_out( const b200_geom* g, const int16_t* const* refs, const b200_picture* pic, int threads, int simd, int16_t* const out[3] )
{
  using clk = std::chrono::steady_clock;
  threads = std::max( 1, threads );
  FakePicture cur( *g, 1 );
  std::unique_ptr<FakePicture> ref[4];
  for( int s = 0; s < 4; s++ )
  {
    ref[s].reset( new FakePicture( *g, 1 ) );
    int16_t* p3[3] = { (int16_t*) refs[s * 3], (int16_t*) refs[s * 3 + 1], (int16_t*) refs[s * 3 + 2] };
    ref[s]->setPlanes( *g, p3 );
    if( s ) ref[s]->pic.extendPicBorder();
  }
  CodingStructure& cs = *cur.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  SPS& sps = *cur.sps;
  sps.setUseBIO( true ); sps.setUseDMVR( true ); sps.setUseBcw( true ); sps.setUseAffine( true ); sps.setUseAffineType( true ); sps.setUsePROF( true );
  sps.setUseALF( true ); sps.setUseCCALF( true );
  cur.pps->setLoopFilterAcrossSlicesEnabledFlag( true ); cur.pps->setLoopFilterAcrossTilesEnabledFlag( true );
  Slice* sl = cur.pic.slices[0];
  sl->setSliceType( B_SLICE ); sl->setPOC( 8 );
  const int pocs[4] = { 4, 0, 12, 16 };
  for( int l = 0; l < 2; l++ ) for( int i = 0; i < 2; i++ )
  { sl->m_apcRefPicList[l][i] = &ref[l * 2 + i]->pic; sl->m_aiRefPOCList[l][i] = pocs[l * 2 + i]; sl->m_bIsUsedAsLongTerm[l][i] = false; ref[l * 2 + i]->pic.poc = pocs[l * 2 + i]; }
  sl->setNumRefIdx( REF_PIC_LIST_0, 2 ); sl->setNumRefIdx( REF_PIC_LIST_1, 2 );
  sl->resetWpScaling();
  applyWp( cur, sl );
  { SliceMap sm; sm.addCtusToSlice( 0, pcv.widthInCtus, 0, pcv.heightInCtus, pcv.widthInCtus ); sl->setSliceMap( sm ); }
  std::vector<MotionInfo> motion( (size_t) pcv.num4x4CtuBlks * pcv.sizeInCtus );
  std::vector<Mv> dmvrCache( std::max<size_t>( pic->numDmvr + 64, (size_t) pcv.num8x8CtuBlks * pcv.sizeInCtus ) );
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) cs.getCtuData( a ).motion = motion.data() + (size_t) a * pcv.num4x4CtuBlks;
  cs.m_dmvrMvCache = dmvrCache.data();
  addCtuCUs( cur );
  if( pic->given[0] ) { int16_t* gp[3] = { (int16_t*) pic->given[0], (int16_t*) pic->given[1], (int16_t*) pic->given[2] }; cur.setPlanes( *g, gp ); }   // pre-reconstructed (intra) samples
  g_tCoeffOps = simd ? g_simdOps : g_scalarOps;

  // deblocking / SAO / ALF parameters
  const bool doLf = pic->flags & B200_PIC_DEBLOCK, doSao = pic->flags & B200_PIC_SAO, doAlf = pic->flags & B200_PIC_ALF;
  if( doLf )
  {
    sl->setDeblockingFilterDisable( pic->lfSlices[0].disable );
    sl->setDeblockingFilterBetaOffsetDiv2( pic->lfSlices[0].betaOffsetDiv2[0] ); sl->setDeblockingFilterTcOffsetDiv2( pic->lfSlices[0].tcOffsetDiv2[0] );
    sl->setDeblockingFilterCbBetaOffsetDiv2( pic->lfSlices[0].betaOffsetDiv2[1] ); sl->setDeblockingFilterCbTcOffsetDiv2( pic->lfSlices[0].tcOffsetDiv2[1] );
    sl->setDeblockingFilterCrBetaOffsetDiv2( pic->lfSlices[0].betaOffsetDiv2[2] ); sl->setDeblockingFilterCrTcOffsetDiv2( pic->lfSlices[0].tcOffsetDiv2[2] );
    cur.setLfGrid( 0, pic->lfV ); cur.setLfGrid( 1, pic->lfH );
  }
  PelStorage fltBuf; fltBuf.create( pcv.chrFormat, Size( g->width, g->height ), g->ctuSize, 16, MEMORY_ALIGN_DEF_SIZE );
  if( doSao ) for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
  {
    SAOBlkParam& bp = cs.getCtuData( a ).saoParam; bp.reset();
    for( int c = 0; c < 3; c++ )
    {
      const b200_sao_ctu& sc = pic->sao[a];
      if( sc.type[c] == B200_SAO_OFF ) continue;
      bp[c].modeIdc = SAO_MODE_NEW; bp[c].typeIdc = sc.type[c]; bp[c].typeAuxInfo = sc.band[c];
      if( sc.type[c] == B200_SAO_BO ) for( int i = 0; i < 4; i++ ) bp[c].offset[( sc.band[c] + i ) & 31] = sc.offset[c][i];
      else for( int i = 0; i < 5; i++ ) bp[c].offset[i] = sc.offset[c][i];
    }
  }
  static std::shared_ptr<APS> apsStore[ALF_CTB_MAX_NUM_APS];
  if( doAlf )
  {
    const b200_alf_tables* T = pic->alfTabs;
    const APS* apss[ALF_CTB_MAX_NUM_APS] = { nullptr };
    for( int i = 0; i < ALF_CTB_MAX_NUM_APS; i++ ) { apsStore[i] = std::make_shared<APS>(); apsStore[i]->setAPSId( i ); apss[i] = apsStore[i].get(); }
    AlfApsIdVec ids;
    for( int i = 0; i < T->numLumaSets - 16; i++ )
    {
      AlfSliceParam& p = apsStore[i]->getAlfAPSParam();
      memcpy( p.lumaCoeffFinal, T->lumaCoeff + (size_t) ( 16 + i ) * 1300, 2600 ); memcpy( p.lumaClippFinal, T->lumaClip + (size_t) ( 16 + i ) * 1300, 2600 );
      p.lumaFinalDone = true; ids.push_back( i );
    }
    sl->setNumAlfAps( T->numLumaSets - 16 ); sl->setAlfApsIdsLuma( ids );
    AlfSliceParam& pc = apsStore[7]->getAlfAPSParam();
    pc.numAlternativesChroma = T->numChromaAlts;
    memcpy( pc.chromaCoeff, T->chromaCoeff, T->numChromaAlts * 14 ); memcpy( pc.chrmClippFinal, T->chromaClip, T->numChromaAlts * 14 ); pc.chrmFinalDone = true;
    sl->setAlfApsIdChroma( 7 );
    for( int c = 0; c < 2; c++ ) for( int k = 0; k < T->numCc[c]; k++ ) memcpy( apsStore[6 - c]->getCcAlfAPSParam().ccAlfCoeff[c][k], T->ccCoeff[c] + k * 7, 14 );
    sl->setCcAlfCbEnabledFlag( T->numCc[0] > 0 ); sl->setCcAlfCrEnabledFlag( T->numCc[1] > 0 ); sl->setCcAlfCbApsId( 6 ); sl->setCcAlfCrApsId( 5 );
    sl->setAlfApss( apss );
    for( int c = 0; c < 3; c++ ) sl->setAlfEnabledFlag( ComponentID( c ), true );
    for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
    {
      CtuAlfData& d = cs.getCtuData( a ).alfParam;
      for( int c = 0; c < 3; c++ ) d.alfCtuEnableFlag[c] = pic->alf[a].enable[c] & 1;
      d.alfCtbFilterIndex = pic->alf[a].lumaSet;
      for( int c = 0; c < 2; c++ ) { d.alfCtuAlternative[c] = pic->alf[a].chromaAlt[c]; d.ccAlfFilterControl[c] = pic->alf[a].ccIdx[c]; }
    }
  }
  // per-thread resources (as DecLibRecon's m_pcThreadResource)
  static RdCost rdScalar( false ), rdSimd( true );
  std::vector<std::unique_ptr<InterPrediction>> ips; std::vector<std::unique_ptr<TrQuant>> tqs; std::vector<std::unique_ptr<Quant>> qns;
  for( int t = 0; t < threads; t++ )
  {
    ips.emplace_back( new InterPrediction() ); ips.back()->init( simd ? &rdSimd : &rdScalar, pcv.chrFormat, g->ctuSize, simd != 0 );
    tqs.emplace_back( new TrQuant( ips.back().get() ) ); qns.emplace_back( new Quant( nullptr, simd != 0 ) );
  }
  const bool doLmcs = ( pic->flags & B200_PIC_LMCS ) && pic->lmcs;
  std::vector<std::unique_ptr<Reshape>> rsps;
  if( doLmcs )
  {
    initPelOps();
    for( int t = 0; t < threads; t++ ) { rsps.emplace_back( new Reshape() ); fillReshapeFromTables( *rsps.back(), g->bitDepth, pic->lmcs ); }
  }
  TCoeffOps ops = simd ? g_simdOps : g_scalarOps;
  LoopFilter lf( simd != 0 );
  SampleAdaptiveOffset sao( simd != 0 );
  AdaptiveLoopFilter alf( simd != 0 );
  if( doAlf ) alf.create( cur.ph.get(), cur.sps.get(), cur.pps.get(), threads, fltBuf );
  PelUnitBuf reco = cs.getRecoBuf();
  const int wC = pcv.widthInCtus, hC = pcv.heightInCtus;

  // ------------------------------------------------------------------ timed region
  WorkerPool pool( threads );
  if( doSao ) sao.create( g->width, g->height, pcv.chrFormat, g->ctuSize, g->ctuSize, 0, 0, fltBuf );
  const auto t0 = clk::now();
  // one reference-picture border extension per picture, as the six tasks DecLibRecon.cpp:291-374 issues
  parallelFor( 6, pool, [&]( int k, int ) {
    Picture& rpic = ref[0]->pic;
    switch( k ) {
      case 0: rpic.extendPicBorder( true, false, false, false ); break;
      case 1: rpic.extendPicBorder( false, true, false, false ); break;
      case 2: rpic.extendPicBorder( false, false, true, false, CH_L ); break;
      case 3: rpic.extendPicBorder( false, false, false, true, CH_L ); break;
      case 4: rpic.extendPicBorder( false, false, true, false, CH_C ); break;
      default: rpic.extendPicBorder( false, false, false, true, CH_C ); break;
    } } );
  // K2: per PU (dynamic scheduling over chunks of 16 PUs)
  {
    const int nChunks = ( (int) pic->numPus + 15 ) / 16;
    parallelFor( nChunks, pool, [&]( int ch, int t ) {
      CodingUnit cu;
      for( size_t n = (size_t) ch * 16; n < std::min<size_t>( pic->numPus, (size_t) ch * 16 + 16 ); n++ )
      {
        const b200_pu& pu = pic->pus[n];
        fillCuFromPu( cu, pu, cur, sl );
        PelUnitBuf predBuf = reco.subBuf( UnitArea( pcv.chrFormat, Area( pu.x, pu.y, pu.w, pu.h ) ) );   // rootCbf==0 style: MC writes straight into the picture (DecCu.cpp:405)
        runMc( *ips[t], cu, predBuf );
        if( doLmcs ) rsps[t]->rspBufFwd( predBuf.Y() );             // DecCu.cpp:458-476
      }
    } );
  }
  // K1: per TU chunk.  With LMCS chroma scaling (DecCu.cpp:483 finishLMCSAndReco) the luma TUs go first: the chroma scale of a VPDU is
  // derived from the reconstructed (mapped-domain) luma next to it — for an all-inter picture that is order-independent once luma is done.
  {
    const bool twoPhase = doLmcs && pic->lmcs->chromaAdj;
    const int nChunks = ( (int) pic->numTus + 31 ) / 32;
    for( int phase = 0; phase < ( twoPhase ? 2 : 1 ); phase++ )
    parallelFor( nChunks, pool, [&]( int ch, int t ) {
      TCoeff* dq  = tqs[t]->m_dqnt; TCoeff* tmp = tqs[t]->m_tmp; TCoeff* blk = tqs[t]->m_blk;
      alignas( 32 ) Pel r0[64 * 64]; alignas( 32 ) Pel r1[64 * 64];
      const int pmax = ( 1 << g->bitDepth ) - 1;
      for( size_t n = (size_t) ch * 32; n < std::min<size_t>( pic->numTus, (size_t) ch * 32 + 32 ); n++ )
      {
        const b200_tu& tu = pic->tus[n];
        if( twoPhase && ( tu.comp != 0 ) != ( phase == 1 ) ) continue;
        const int w = 1 << tu.log2w, h = 1 << tu.log2h;
        refTuResidual( tu, g->bitDepth, pic->coefs, *tqs[t], *qns[t], ops, dq, tmp, blk, r0, w );
        int nOut = 1, comp1 = 0;
        if( tu.ict ) { comp1 = tu.comp == 1 ? 2 : 1; nOut = 2; const int m = tu.ict; for( int i = 0; i < w * h; i++ ) { const int c = r0[i]; r1[i] = Pel( m == 2 ? c : m == -2 ? -c : m > 0 ? ( c >> 1 ) : ( ( -c ) >> 1 ) ); } }
        if( twoPhase && tu.comp != 0 && w * h > 4 )
        {
          // the real Reshape::calculateChromaAdjVpduNei on the (one CU per CTU) structure + the real AreaBuf::scaleSignal
          TransformUnit rtu; memset( (void*) &rtu, 0, sizeof( rtu ) );
          const Position lumaPos( tu.x * 2, tu.y * 2 );
          rtu.cu = cs.getCU( lumaPos, CH_L );
          rsps[t]->setVPDULoc( -1, -1 );
          const int sc = rsps[t]->calculateChromaAdjVpduNei( rtu, lumaPos );
          ClpRng rng; rng.bd = g->bitDepth;
          PelBuf b0( r0, w, w, h ); b0.scaleSignal( sc, rng );
          if( nOut == 2 ) { PelBuf b1( r1, w, w, h ); b1.scaleSignal( sc, rng ); }
        }
        for( int o = 0; o < nOut; o++ )
        {
          PelBuf pb = reco.bufs[o ? comp1 : tu.comp]; const Pel* r = o ? r1 : r0;
          Pel* d = pb.buf + tu.y * pb.stride + tu.x;
          for( int y = 0; y < h; y++, d += pb.stride, r += w ) for( int x = 0; x < w; x++ ) d[x] = Pel( Clip3( 0, pmax, d[x] + r[x] ) );   // recoCore, Buffer.cpp:83
        }
      }
    } );
  }
  if( doLmcs )   // RSP stage (DecLibRecon.cpp:935): inverse map per CTU
    parallelFor( wC * hC, pool, [&]( int a, int t ) {
      const int x = ( a % wC ) * g->ctuSize, y = ( a / wC ) * g->ctuSize, w = std::min( g->ctuSize, g->width - x ), h = std::min( g->ctuSize, g->height - y );
      PelBuf b = reco.bufs[0].subBuf( Position( x, y ), Size( w, h ) );
      const Reshape& r = *rsps[t];
      const PelBufferOps& pops = simd ? g_pelOpsSimd : g_pelOpsScalar;
      if( pops.rspBcw ) pops.rspBcw( b.buf, b.stride, w, h, r.m_lumaBD, r.m_sliceReshapeInfo.reshaperModelMinBinIdx, r.m_sliceReshapeInfo.reshaperModelMaxBinIdx, r.m_reshapePivot.data(), r.m_invScaleCoef.data(), r.m_inputPivot.data() );
      else              pops.applyLut( b.buf, b.stride, w, h, r.m_invLUT );
    } );
  if( doLf )
  {
    parallelFor( wC * hC, pool, [&]( int a, int ) { lf.loopFilterCTU( cs, MAX_NUM_CHANNEL_TYPE, a % wC, a / wC, EDGE_VER ); } );
    parallelFor( wC * hC, pool, [&]( int a, int ) { lf.loopFilterCTU( cs, MAX_NUM_CHANNEL_TYPE, a % wC, a / wC, EDGE_HOR ); } );
  }
  if( doSao )
  {
    // SAOPrepareCTULine copies the deblocked rows (SampleAdaptiveOffset.cpp:400), one CTU line per task
    parallelFor( hC, pool, [&]( int r, int ) {
      const UnitArea line = clipArea( UnitArea( pcv.chrFormat, Area( 0, r * g->ctuSize, g->width, g->ctuSize ) ), cur.pic );
      fltBuf.subBuf( line ).copyFrom( cs.getRecoBuf().subBuf( line ) ); } );
    parallelFor( wC * hC, pool, [&]( int a, int ) {
      sao.SAOProcessCTU( cs, clipArea( UnitArea( pcv.chrFormat, Area( ( a % wC ) * g->ctuSize, ( a / wC ) * g->ctuSize, g->ctuSize, g->ctuSize ) ), cur.pic ) ); } );
  }
  if( doAlf )
  {
    parallelFor( wC * hC, pool, [&]( int a, int ) { alf.prepareCTU( cs, a % wC, a / wC ); } );
    parallelFor( wC * hC, pool, [&]( int a, int t ) { alf.processCTU( cs, a % wC, a / wC, t ); } );
  }
  const double secs = std::chrono::duration<double>( clk::now() - t0 ).count();
  if( out ) { if( doAlf ) cur.getPlanes( *g, out, true, &fltBuf ); else cur.getPlanes( *g, out ); }
  return secs;
}

extern "C" double ref_decompress_picture_mt( const b200_geom* g, const int16_t* const* refs, const b200_picture* pic, int threads, int simd )
{
  return ref_decompress_picture_out( g, refs, pic, threads, simd, nullptr );
}

// ================================================================================================ the DecLibRecon seam, executed
#include "ref_seam.h"
