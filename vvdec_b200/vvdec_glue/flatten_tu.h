// flatten_tu.h — host-side flattener, TU part: VVdeC TransformUnit -> b200_tu record (+ packed levels).
//
// This is the glue that lives INSIDE a VVdeC build (it includes the reference's private headers, like
// the reference's own unit test does, tests/vvdec_unit_test/CMakeLists.txt:27-38) and feeds the C ABI
// in include/vvdec_b200.h.  It contains no pixel arithmetic: it only derives the per-TU scalars that
// Quant::dequant (Quant.cpp:295-349), TrQuant::xInvLfnst (TrQuant.cpp:201-232) and TrQuant::getTrTypes
// (TrQuant.cpp:330) derive on the CPU before they call the function-pointer kernels.
#pragma once
#include <vector>
#include "vvdec_b200.h"
#include "CommonLib/CommonDef.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/Slice.h"
#include "CommonLib/Quant.h"
#include "CommonLib/TrQuant.h"
#include "CommonLib/Rom.h"

namespace b200glue
{
using namespace vvdec;

// Which component carries the coded levels of a joint-CbCr TU (DecCu.cpp:552-568).
inline ComponentID jccrCodedComp( const TransformUnit& tu ) { return ( tu.jointCbCr >> 1 ) ? COMPONENT_Cb : COMPONENT_Cr; }

// Fills one record for (tu, compID) and appends its level corner to `coefs`.
// `levels` is the reco-plane area of the TU component, which holds the parsed int16 levels in place
// (CABACReader.cpp:2457-2478).  Returns false if the component carries no residual.
inline bool flattenTU( const TransformUnit& tu, const ComponentID compID, TrQuant& trq, std::vector<int16_t>& coefs, b200_tu& r )
{
  const CodingUnit& cu  = *tu.cu;
  const SPS&        sps = *cu.sps;
  const CompArea&   area = tu.blocks[compID];
  if( !area.valid() ) return false;

  int ict = 0;
  if( tu.jointCbCr && isChroma( compID ) )
  {
    if( compID != jccrCodedComp( tu ) ) return false;             // the other plane is derived by the kernel
    ict = TU::getICTMode( tu, cu.cs->picHeader->getJointCbCrSignFlag() );
  }
  else if( !TU::getCbf( tu, compID ) ) return false;

  memset( &r, 0, sizeof( r ) );
  r.x = area.x; r.y = area.y;
  r.log2w = getLog2( area.width ); r.log2h = getLog2( area.height );
  r.comp = compID;
  r.ict  = (int8_t) ict;

  const bool isTS    = tu.mtsIdx( compID ) == MTS_SKIP;
  const int  bdpcm   = isLuma( compID ) ? cu.bdpcmMode() : cu.bdpcmModeChroma();
  if( isTS )       r.flags |= B200_TU_TS;
  if( bdpcm == 1 ) r.flags |= B200_TU_BDPCM_H;
  if( bdpcm == 2 ) r.flags |= B200_TU_BDPCM_V;

  // ---- Quant::dequant scalars (Quant.cpp:295-349) ----
  const QpParam cQP( tu, compID );
  const int  maxLog2TrDynamicRange = sps.getMaxLog2TrDynamicRange( toChannelType( compID ) );
  const bool lfnstApplied = cu.lfnstIdx() > 0 && ( CU::isSepTree( cu ) ? true : isLuma( compID ) );
  const bool explicitSL   = cu.slice->getExplicitScalingListUsed();
  const bool disableSMForLFNST = explicitSL ? sps.getDisableScalingMatrixForLfnstBlks() : false;
  const bool disableSMForACT   = sps.getScalingMatrixForAlternativeColourSpaceDisabledFlag() && sps.getScalingMatrixDesignatedColourSpaceFlag() == cu.colorTransform();
  const bool enableSL = explicitSL && !( isTS || ( lfnstApplied && disableSMForLFNST ) || disableSMForACT );   // Quant.h getUseScalingList
  const int  trShift  = maxLog2TrDynamicRange - sps.getBitDepth() - ( ( r.log2w + r.log2h ) >> 1 );
  const bool sqrt2    = TU::needsBlockSizeTrafoScale( tu, compID );
  const int  iTransformShift = trShift + ( sqrt2 ? -1 : 0 );
  const bool depQuant = cu.slice->getDepQuantEnabledFlag() && !isTS;
  const int  qpPer    = depQuant ? ( ( cQP.Qp( isTS ) + 1 ) / 6 ) : cQP.per( isTS );
  const int  qpRem    = depQuant ? ( cQP.Qp( isTS ) + 1 - 6 * qpPer ) : cQP.rem( isTS );
  const int  rightShift = IQUANT_SHIFT + ( depQuant ? 1 : 0 ) - ( ( isTS ? 0 : iTransformShift ) + qpPer ) + ( enableSL ? LOG2_SCALING_LIST_NEUTRAL_VALUE : 0 );
  const int  scaleBits  = IQUANT_SHIFT + 1;
  const uint32_t inBits = std::min<uint32_t>( maxLog2TrDynamicRange + 1, (int) sizeof( Intermediate_Int ) * 8 + rightShift - scaleBits );
  r.rightShift = (int8_t) rightShift;
  r.inBits     = (uint8_t) inBits;
  r.scale      = (uint8_t) g_InvQuantScales[sqrt2 ? 1 : 0][qpRem];
  if( enableSL ) r.flags |= B200_TU_SCALING;    // caller sets slOff from its per-picture scaling arena

  const int maxX = bdpcm ? area.width  - 1 : tu.maxScanPosX[compID];
  const int maxY = bdpcm ? area.height - 1 : tu.maxScanPosY[compID];
  r.maxX = (uint8_t) maxX; r.maxY = (uint8_t) maxY;

  // ---- transform types (TrQuant.cpp:330) ----
  int trH = DCT2, trV = DCT2;
  if( !isTS ) trq.getTrTypes( tu, compID, trH, trV );
  r.trType = (uint8_t) ( trH | ( trV << 2 ) );

  // ---- LFNST side info (TrQuant.cpp:201-232) ----
  if( sps.getUseLFNST() && cu.lfnstIdx() && !isTS && lfnstApplied )
  {
    uint32_t intraMode;
    if( CU::isMIP( cu, toChannelType( compID ) ) ) intraMode = PLANAR_IDX;
    else intraMode = PU::isLMCMode( cu.intraDir[toChannelType( compID )] ) ? PU::getCoLocatedIntraLumaMode( cu ) : PU::getFinalIntraMode( cu, toChannelType( compID ) );
    intraMode = trq.getLFNSTIntraMode( PU::getWideAngIntraMode( tu, intraMode, compID ) );
    const bool transpose = trq.getTransposeFlag( intraMode );
    r.lfnst = (uint8_t) ( cu.lfnstIdx() | ( g_lfnstLut[intraMode] << 2 ) | ( transpose ? 16 : 0 ) );
  }

  // ---- pack the coded corner ----
  const CPelBuf levels = cu.cs->getRecoBuf( area );
  r.coefOff = (uint32_t) coefs.size();
  for( int y = 0; y <= maxY; y++ )
    for( int x = 0; x <= maxX; x++ )
      coefs.push_back( levels.buf[x + y * levels.stride] );
  return true;
}

}   // namespace b200glue
