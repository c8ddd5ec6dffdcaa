// flatten_intra.h — host-side flattener for K6: one intra TransformUnit component -> b200_intra_tu.  Glue that lives INSIDE a VVdeC build.
// The record carries the decisions DecCu::predAndReco / IntraPrediction take from the CU tree: final mode, multi reference line, whether
// the [1 2 1] reference filter applies, and the neighbour availability xFillReferenceSamples derives (IntraPrediction.cpp:1098-1130) — the
// cs.getCURestricted walks are pointer chasing over the CU map and stay on the host.  No pixel arithmetic.
// Pinned by tests/test_intra_oracle_vs_ref.py (availability against IntraPrediction::m_neighborSize, the rest through the prediction).
#pragma once
#include <string.h>
#include <algorithm>
#include "vvdec_b200.h"
#include "CommonLib/CommonDef.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/IntraPrediction.h"

namespace b200glue
{
using namespace vvdec;

enum FlattenIntraResult { FLATTEN_INTRA_OK = 0, FLATTEN_INTRA_UNSUPPORTED = 1 };

// the TU of `cu` that covers `pos` (IntraPrediction.cpp:1329, file-static there)
inline const TransformUnit* intraTuAt( const CodingUnit& cu, const Position& pos, const ChannelType chType )
{
  const TransformUnit* ptu = &cu.firstTU;
  if( !ptu->next ) return ptu;
  while( !( ptu->blocks[chType].x + ptu->blocks[chType].width > pos.x && ptu->blocks[chType].y + ptu->blocks[chType].height > pos.y ) ) ptu = ptu->next;
  return ptu;
}

// units available along the above(-right) row / left(-below) column starting at posLT: isAboveAvailable / isLeftAvailable (:1343, :1373)
inline int intraUnitsAvailable( const TransformUnit& tu, const ChannelType chType, const Position& posLT, const int numUnits, const int unitSize, const bool above )
{
  const CodingUnit&      cu = *tu.cu;
  const CodingStructure& cs = *cu.cs;
  const int maxD = numUnits * unitSize;
  Position refPos = above ? posLT.offset( 0, -1 ) : posLT.offset( -1, 0 );
  const TransformUnit* nb = nullptr;
  int d = 0;
  while( d < maxD )
  {
    const CodingUnit* cuN = cs.getCURestricted( refPos, cu, chType, nb ? nullptr : ( above ? cu.above : cu.left ) );
    if( !cuN ) break;
    nb = intraTuAt( *cuN, refPos, chType );
    if( cuN->ctuData == cu.ctuData && nb->idx >= tu.idx ) break;
    const int diff = above ? (int) nb->blocks[chType].width - refPos.x + nb->blocks[chType].x : (int) nb->blocks[chType].height - refPos.y + nb->blocks[chType].y;
    d += diff;
    if( above ) refPos.x += diff; else refPos.y += diff;
  }
  return std::min( d / unitSize, numUnits );
}

// wholeCu: the block is the CU's (cu.blocks[compID]) although the CU carries several transform units — CIIP predicts the CU block in one piece also where the
// CU is larger than the maximum transform size (predBlendIntraCiip: initIntraPatternChType( cu.firstTU, cu.blocks[compID] ), IntraPrediction.cpp:909).
inline FlattenIntraResult flattenIntraTU( const TransformUnit& tu, const ComponentID compID, b200_intra_tu& r, const bool wholeCu = false )
{
  const CodingUnit&      cu     = *tu.cu;
  const CodingStructure& cs     = *cu.cs;
  const PreCalcValues&   pcv    = *cs.pcv;
  const ChannelType      chType = toChannelType( compID );
  const CompArea&        area   = wholeCu ? cu.blocks[compID] : tu.blocks[compID];
  memset( &r, 0, sizeof( r ) );
  if( cu.colorTransform() || ( isLuma( compID ) && cu.ispMode() ) ) return FLATTEN_INTRA_UNSUPPORTED;
  const bool mip = CU::isMIP( cu, chType );
  if( mip && !isLuma( compID ) ) return FLATTEN_INTRA_UNSUPPORTED;                                       // MIP chroma exists in 4:4:4 only
  const uint32_t finalMode = PU::getFinalIntraMode( cu, chType );
  const bool lm = !isLuma( compID ) && PU::isLMCMode( finalMode );
  if( lm && cu.chromaFormat != CHROMA_420 ) return FLATTEN_INTRA_UNSUPPORTED;
  const int bdpcm = isLuma( compID ) ? cu.bdpcmMode() : cu.bdpcmModeChroma();
  r.x = (uint16_t) area.x; r.y = (uint16_t) area.y; r.log2w = (uint8_t) getLog2( area.width ); r.log2h = (uint8_t) getLog2( area.height ); r.comp = (uint8_t) compID;
  r.mode = bdpcm ? ( bdpcm == 1 ? B200_INTRA_BDPCM_HOR : B200_INTRA_BDPCM_VER ) : (uint8_t) finalMode;
  if( lm ) r.mode = (uint8_t) ( B200_INTRA_LM + ( finalMode - LM_CHROMA_IDX ) );
  if( mip ) { r.mode = B200_INTRA_MIP; r.mip = (uint8_t) ( cu.intraDir[CHANNEL_TYPE_LUMA] | ( cu.mipTransposedFlag() ? 0x80 : 0 ) ); }        // predIntraMip :1926
  r.multiRefIdx = isLuma( compID ) ? (uint8_t) cu.multiRefIdx() : 0;
  if( !mip && isLuma( compID ) && IntraPrediction::useFilteredIntraRefSamples( compID, cu, tu ) ) r.flags |= B200_INTRA_FILTER_REF;    // DecCu.cpp:339 (MIP: unfiltered, :323)

  // neighbourhood (xFillReferenceSamples :1086-1130).  The reference analyses it once per TU, for the first component of the CU's channel
  // type, and reuses the three counts for the other components (m_lastCUidx, :1101): in a single tree the chroma blocks take the luma block's.
  // ISP CUs: the luma of the first sub-partition analysed the whole CU (initIntraPatternChTypeISP :997, with cu.firstTU), and that is what the chroma blocks in the
  // last transform unit inherit.
  const bool        isp     = cu.ispMode() != 0 || wholeCu;                                  // (the neighbourhood of the CU block, analysed through cu.firstTU)
  const TransformUnit& tuA  = isp ? cu.firstTU : tu;
  const ComponentID anaComp = getFirstComponentOfChannel( cu.chType() );
  const ChannelType anaCh   = toChannelType( anaComp );
  const CompArea&   ana     = isp ? cu.blocks[anaComp] : tu.blocks[anaComp].valid() ? tu.blocks[anaComp] : area;
  const ChannelType ch      = ( isp || tu.blocks[anaComp].valid() ) ? anaCh : chType;
  const int csx = getChannelTypeScaleX( ch, pcv.chrFormat ), csy = getChannelTypeScaleY( ch, pcv.chrFormat );
  const int unitW = pcv.minCUWidth >> csx, unitH = pcv.minCUHeight >> csy;
  const int totalAbove = ( 2 * (int) ana.width + unitW - 1 ) / unitW, totalLeft = ( 2 * (int) ana.height + unitH - 1 ) / unitH;
  const int numAbove = ana.width / unitW, numLeft = ana.height / unitH;
  const Position posLT = ana.pos();
  const bool sameCTU = ( posLT.x & ( pcv.maxCUWidthMask >> csx ) ) && ( posLT.y & ( pcv.maxCUHeightMask >> csy ) );
  if( sameCTU || cs.getCURestricted( posLT.offset( -1, -1 ), cu, ch, cu.left ? cu.left : cu.above ) ) r.flags |= B200_INTRA_AVAIL_TL;
  if( cu.above || ana.y > cu.blocks[ch].y )
    r.numAbove = (uint8_t) ( numAbove + intraUnitsAvailable( tuA, ch, Position( posLT.x + (PosType) ana.width, posLT.y ), totalAbove - numAbove, unitW, true ) );
  if( cu.left || ana.x > cu.blocks[ch].x )
    r.numLeft = (uint8_t) ( numLeft + intraUnitsAvailable( tuA, ch, Position( posLT.x, posLT.y + (PosType) ana.height ), totalLeft - numLeft, unitH, false ) );
  if( lm )
  {
    // CCLM: what xGetLumaRecPixels (:1461-1465) and xGetLMParameters (:1762-1795) derive from the CU map, in the chroma channel
    const int unit = ( 1 << MIN_CU_LOG2 ) >> getComponentScaleX( compID, cu.chromaFormat );
    const bool above = cu.above || area.y > cu.blocks[CH_C].y, left = cu.left || area.x > cu.blocks[CH_C].x;
    if( above ) r.flags |= B200_INTRA_LM_ABOVE;
    if( left )  r.flags |= B200_INTRA_LM_LEFT;
    if( cu.sps->getCclmCollocatedChromaFlag() ) r.flags |= B200_INTRA_LM_COLLOCATED;
    const int aboveUnits = area.width / unit, leftUnits = area.height / unit;
    const int totalAbove = ( 2 * (int) area.width + unit - 1 ) / unit, totalLeft = ( 2 * (int) area.height + unit - 1 ) / unit;
    r.lmAbove = above ? (uint8_t) aboveUnits : 0; r.lmLeft = left ? (uint8_t) leftUnits : 0;
    if( finalMode == MDLM_T_IDX && above )
      r.lmAbove = (uint8_t) ( aboveUnits + intraUnitsAvailable( tu, CHANNEL_TYPE_CHROMA, Position( area.x + (PosType) area.width, area.y ), std::min( totalAbove - aboveUnits, (int) area.height / unit ), unit, true ) );
    if( finalMode == MDLM_L_IDX && left )
      r.lmLeft = (uint8_t) ( leftUnits + intraUnitsAvailable( tu, CHANNEL_TYPE_CHROMA, Position( area.x, area.y + (PosType) area.height ), std::min( totalLeft - leftUnits, (int) area.width / unit ), unit, false ) );
  }
  return FLATTEN_INTRA_OK;
}

// ISP CU: one B200_INTRA_ISP record per luma prediction region (include/vvdec_b200.h), in decoding order.  The CU-level neighbourhood is what
// initIntraPatternChTypeISP (IntraPrediction.cpp:966) derives with its first call; the per-region availability of the CU's left / above neighbour (:971-972)
// is the CU's.  emit( const b200_intra_tu& ) is called per region.
template<class Emit>
inline FlattenIntraResult flattenIspCu( const CodingUnit& cu, Emit emit )
{
  if( !cu.ispMode() || !isLuma( cu.chType() ) || cu.colorTransform() || cu.mipFlag() || cu.multiRefIdx() || cu.bdpcmMode() ) return FLATTEN_INTRA_UNSUPPORTED;
  const CodingStructure& cs = *cu.cs; const PreCalcValues& pcv = *cs.pcv;
  const CompArea& Y = cu.Y();
  const int W = Y.width, H = Y.height;
  const bool hor = cu.ispMode() == HOR_INTRA_SUBPARTITIONS;
  const int part = (int) CU::getISPSplitDim( W, H, hor ? TU_1D_HORZ_SPLIT : TU_1D_VERT_SPLIT );
  const bool regDiff = CU::isPredRegDiffFromTB( cu, COMPONENT_Y );                  // vertical split with sub-partitions narrower than 4
  const int rw = hor ? W : ( regDiff ? 4 : part ), rh = hor ? part : H, nReg = hor ? H / part : W / rw;
  b200_intra_tu base; memset( &base, 0, sizeof( base ) );
  base.comp = 0; base.mode = (uint8_t) PU::getFinalIntraMode( cu, CH_L ); base.flags = B200_INTRA_ISP;
  base.log2w = (uint8_t) getLog2( rw ); base.log2h = (uint8_t) getLog2( rh );
  // CU-level neighbourhood (xFillReferenceSamples :1098-1130 on cu.Y() with cu.firstTU)
  const int unit = pcv.minCUWidth, totalAbove = ( 2 * W + unit - 1 ) / unit, totalLeft = ( 2 * H + unit - 1 ) / unit, numAbove = W / unit, numLeft = H / unit;
  const Position posLT = Y.pos();
  const bool sameCTU = ( posLT.x & pcv.maxCUWidthMask ) && ( posLT.y & pcv.maxCUHeightMask );
  if( sameCTU || cs.getCURestricted( posLT.offset( -1, -1 ), cu, CH_L, cu.left ? cu.left : cu.above ) ) base.flags |= B200_INTRA_AVAIL_TL;
  if( cu.above ) base.numAbove = (uint8_t) ( numAbove + intraUnitsAvailable( cu.firstTU, CH_L, Position( posLT.x + W, posLT.y ), totalAbove - numAbove, unit, true ) );
  if( cu.left )  base.numLeft  = (uint8_t) ( numLeft  + intraUnitsAvailable( cu.firstTU, CH_L, Position( posLT.x, posLT.y + H ), totalLeft - numLeft, unit, false ) );
  base.lmLeft  = nullptr != cs.getCURestricted( posLT.offset( -1, 0 ), cu, CH_L, cu.left );
  base.lmAbove = nullptr != cs.getCURestricted( posLT.offset( 0, -1 ), cu, CH_L, cu.above );
  // residual flags of the transform units, region by region
  const TransformUnit* tu = &cu.firstTU;
  const int tusPerReg = regDiff ? rw / part : 1;
  for( int k = 0; k < nReg; k++ )
  {
    b200_intra_tu r = base;
    r.x = (uint16_t) ( Y.x + ( hor ? 0 : k * rw ) ); r.y = (uint16_t) ( Y.y + ( hor ? k * rh : 0 ) );
    r.mip = (uint8_t) ( cu.ispMode() | ( k << 2 ) | ( getLog2( nReg ) << 4 ) );
    for( int i = 0; i < tusPerReg; i++, tu = tu->next )
    {
      if( !tu ) return FLATTEN_INTRA_UNSUPPORTED;
      if( TU::getCbf( *tu, COMPONENT_Y ) ) { r.ciip |= (uint8_t) ( 1 << i ); r.flags |= B200_INTRA_ADD_RESI; }
    }
    emit( r );
  }
  return FLATTEN_INTRA_OK;
}

// CIIP CU (inter CU with cu.ciipFlag()): predBlendIntraCiip (IntraPrediction.cpp:887) predicts each component's CU block planar (through
// cu.firstTU, reference filter decided as for an intra CU) and blends it with the inter prediction; the weight follows from whether the left
// (at the bottom-left corner) and above (at the top-right corner) neighbour CUs are intra (:917-925).  Chroma only if it is wider than 2.
// Returns FLATTEN_INTRA_UNSUPPORTED for a component that carries no CIIP block.
inline FlattenIntraResult flattenCiipBlock( const CodingUnit& cu, const ComponentID compID, b200_intra_tu& r )
{
  if( !isLuma( compID ) && !( isChromaEnabled( cu.chromaFormat ) && cu.chromaSize().width > 2 ) ) return FLATTEN_INTRA_UNSUPPORTED;
  if( flattenIntraTU( cu.firstTU, compID, r, true ) != FLATTEN_INTRA_OK ) return FLATTEN_INTRA_UNSUPPORTED;
  r.mode = B200_INTRA_PLANAR; r.multiRefIdx = 0; r.mip = 0;
  r.flags &= ~B200_INTRA_FILTER_REF;
  if( isLuma( compID ) && IntraPrediction::useFilteredIntraRefSamples( COMPONENT_Y, cu, cu ) ) r.flags |= B200_INTRA_FILTER_REF;
  const CodingUnit* cuLeft  = cu.cs->getCURestricted( cu.Y().bottomLeft().offset( -1, 0 ), cu, CHANNEL_TYPE_LUMA, cu.left );
  const CodingUnit* cuAbove = cu.cs->getCURestricted( cu.Y().topRight().offset( 0, -1 ), cu, CHANNEL_TYPE_LUMA, cu.above );
  r.ciip = (uint8_t) ( 3 - !( cuLeft && CU::isIntra( *cuLeft ) ) - !( cuAbove && CU::isIntra( *cuAbove ) ) );
  return FLATTEN_INTRA_OK;
}

}   // namespace b200glue
