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
