// flatten_pu.h — host-side flattener, inter part: VVdeC CodingUnit -> b200_pu record.
//
// Glue that lives INSIDE a VVdeC build (includes the reference's private headers).  It carries no pixel arithmetic: it takes the
// mode decisions that InterPrediction::motionCompensation (InterPrediction.cpp:1372-1442) takes before it calls the kernels, with the
// reference's own predicates (PU::isBiPredFromDifferentDirEqDistPoc, PU::checkDMVRCondition), and writes them into the record's
// flags.  tests/test_flatten_pu_vs_ref.py builds real CodingUnits, runs the real motionCompensation on them and checks that the
// oracle on the flattened record gives the same samples.
#pragma once
#include "vvdec_b200.h"
#include "CommonLib/CommonDef.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/Slice.h"
#include "CommonLib/Picture.h"
#include "CommonLib/Rom.h"

namespace b200glue
{
using namespace vvdec;

// DPB slot of slice->getRefPic( list, refIdx ), supplied by the owner of the device DPB
struct SlotMap { int8_t slot[2][MAX_NUM_REF]; };

enum FlattenPuResult { FLATTEN_PU_OK = 0, FLATTEN_PU_NOT_INTER, FLATTEN_PU_UNSUPPORTED };

// InterPrediction::xCheckIdenticalMotion (:404): both lists point at the same picture with the same motion -> list 0 only
inline bool identicalMotion( const CodingUnit& cu )
{
  const Slice& slice = *cu.slice;
  if( !slice.isInterB() || cu.pps->getWPBiPred() || cu.refIdx[0] < 0 || cu.refIdx[1] < 0 ) return false;
  if( slice.getRefPOC( REF_PIC_LIST_0, cu.refIdx[0] ) != slice.getRefPOC( REF_PIC_LIST_1, cu.refIdx[1] ) ) return false;
  if( !cu.affineFlag() ) return cu.mv[0][0] == cu.mv[1][0];
  return cu.mv[0][0] == cu.mv[1][0] && cu.mv[0][1] == cu.mv[1][1] && ( cu.affineType() == AFFINEMODEL_4PARAM || cu.mv[0][2] == cu.mv[1][2] );
}

// wpIdxOf( refIdx0, refIdx1 ): 1-based index of the b200_wp entry for the combination, 0 if explicit weighting does not apply.
// subPuMC: the CU is one merged run of an SbTMVP CU (m_subPuMC in motionCompensation: no BDOF, no DMVR).
template<class WpIdxOf>
inline FlattenPuResult flattenPU( const CodingUnit& cu, const SlotMap& sm, WpIdxOf wpIdxOf, b200_pu& r, const bool subPuMC = false )
{
  if( !CU::isInter( cu ) || CU::isIBC( cu ) ) return FLATTEN_PU_NOT_INTER;
  // tools the device path does not have: the caller sends such CUs (or the picture) down the CPU path
  // (CIIP CUs: this record is their inter half, BDOF excluded below as in motionCompensation :1412; the intra half is flattenCiipBlock's)
  if( cu.mergeType() == MRG_TYPE_SUBPU_ATMVP || cu.sps->getUseWrapAround() ) return FLATTEN_PU_UNSUPPORTED;   // SbTMVP: flattenSbTmvp
  const Slice& slice = *cu.slice;
  const PPS&   pps   = *cu.pps;
  if( cu.geoFlag() )
  {
    // motionCompensationGeo (:1461-1497): partition p = uni-prediction from list (geoDir_p >> 4) - 1, refIdx geoDir_p & 15, MV cu.mv[p][1]
    memset( &r, 0, sizeof( r ) );
    r.x = cu.lx(); r.y = cu.ly(); r.w = cu.lwidth(); r.h = cu.lheight(); r.interDir = 3; r.dmvrOff = cu.mvdL0SubPuOff;
    const uint8_t dir[2] = { cu.interDirrefIdxGeo0(), cu.interDirrefIdxGeo1() };
    for( int p = 0; p < 2; p++ )
    {
      const int list = ( dir[p] >> 4 ) - 1, idx = dir[p] & 15;
      if( list < 0 || list > 1 || slice.getRefPic( RefPicList( list ), idx )->isRefScaled( cu.pps ) ) return FLATTEN_PU_UNSUPPORTED;
      r.refSlot[p] = sm.slot[list][idx];
      r.mv[p][0] = cu.mv[p][1].hor; r.mv[p][1] = cu.mv[p][1].ver;
    }
    r.flags = B200_PU_GEO | ( cu.imv() == IMV_HPEL ? B200_PU_ALTHPEL : 0 );
    r.bcwW1 = (int8_t) cu.geoSplitDir;
    return FLATTEN_PU_OK;
  }
  for( int l = 0; l < 2; l++ )
    if( cu.refIdx[l] >= 0 && slice.getRefPic( RefPicList( l ), cu.refIdx[l] )->isRefScaled( cu.pps ) ) return FLATTEN_PU_UNSUPPORTED;   // RPR
  if( slice.getRefPic( REF_PIC_LIST_0, 0 )->subPictures.size() > 1 ) return FLATTEN_PU_UNSUPPORTED;                                   // clipMvInSubpic

  memset( &r, 0, sizeof( r ) );
  r.x = cu.lx(); r.y = cu.ly(); r.w = cu.lwidth(); r.h = cu.lheight();
  r.interDir = cu.interDir();
  r.dmvrOff  = cu.mvdL0SubPuOff;
  int refIdx[2] = { cu.refIdx[0], cu.refIdx[1] };

  // ---- the decisions of motionCompensation (:1402-1442) ----
  const WPScalingParam *wp0 = nullptr, *wp1 = nullptr;
  slice.getWpScaling( REF_PIC_LIST_0, refIdx[0], wp0 );
  slice.getWpScaling( REF_PIC_LIST_1, refIdx[1], wp1 );
  bool bioApplied = false;
  if( cu.sps->getUseBIO() && !cu.cs->picHeader->getDisBdofFlag() && !subPuMC &&
      !( cu.affineFlag() || cu.ciipFlag() || cu.smvdMode() || ( cu.sps->getUseBcw() && cu.BcwIdx() != BCW_DEFAULT ) ) )
  {
    const bool wpPresent = wp0[0].bPresentFlag || wp0[1].bPresentFlag || wp0[2].bPresentFlag || wp1[0].bPresentFlag || wp1[1].bPresentFlag || wp1[2].bPresentFlag;
    const bool biocheck0 = !( wpPresent && slice.getSliceType() == B_SLICE );
    const bool biocheck1 = !( pps.getUseWP() && slice.getSliceType() == P_SLICE );
    bioApplied = biocheck0 && biocheck1 && PU::isBiPredFromDifferentDirEqDistPoc( cu ) && cu.Y().height >= 8 && cu.Y().width >= 8 && cu.Y().area() >= 128;
  }
  const bool dmvrApplied = !subPuMC && PU::checkDMVRCondition( cu );
  if( !bioApplied && !dmvrApplied && identicalMotion( cu ) ) refIdx[1] = -1;   // xPredInterUni( list 0 ) only (:1433); interDir stays the CU's (affine spread test)

  for( int l = 0; l < 2; l++ )
  {
    r.refSlot[l] = refIdx[l] < 0 ? -1 : sm.slot[l][refIdx[l]];
    r.mv[l][0] = cu.mv[l][0].hor; r.mv[l][1] = cu.mv[l][0].ver;
    r.cpmv[l][0][0] = cu.mv[l][1].hor; r.cpmv[l][0][1] = cu.mv[l][1].ver;
    r.cpmv[l][1][0] = cu.mv[l][2].hor; r.cpmv[l][1][1] = cu.mv[l][2].ver;
  }
  const bool bi = refIdx[0] >= 0 && refIdx[1] >= 0;

  uint8_t flags = 0;
  if( bioApplied )  flags |= B200_PU_BDOF;
  if( dmvrApplied ) flags |= B200_PU_DMVR;
  if( cu.imv() == IMV_HPEL && !cu.affineFlag() ) flags |= B200_PU_ALTHPEL;                          // xPredInterBlk: useAltHpelIf (:771)
  if( cu.affineFlag() )
  {
    flags |= B200_PU_AFFINE;
    if( cu.affineType() == AFFINEMODEL_6PARAM ) flags |= B200_PU_AFFINE6;
    // xPredAffineBlk :1024-1025: the picture-level part of enablePROF (the CPMV-equality and spread tests are evaluated on the device)
    if( cu.sps->getUsePROF() && !cu.cs->picHeader->getDisProfFlag() ) flags |= B200_PU_PROF0 | B200_PU_PROF1;
  }
  r.flags = flags;

  // ---- bi-prediction combine (xPredInterBi :733-745, xWeightedAverage :1346) ----
  r.bcwW1 = 4;
  r.wpIdx = 0;
  const bool wpB = pps.getWPBiPred() && slice.getSliceType() == B_SLICE && cu.BcwIdx() == BCW_DEFAULT;
  const bool wpP = pps.getUseWP() && slice.getSliceType() == P_SLICE;
  if( !bioApplied && !dmvrApplied && ( wpB || wpP ) ) r.wpIdx = (uint8_t) wpIdxOf( refIdx[0], refIdx[1] );
  else if( bi && cu.BcwIdx() != BCW_DEFAULT && !cu.ciipFlag() ) r.bcwW1 = getBcwWeight( g_BcwInternBcw[cu.BcwIdx()], REF_PIC_LIST_1 );   // xWeightedAverage :1353: a CIIP CU keeps the merge candidate's BcwIdx but averages plainly
  return FLATTEN_PU_OK;
}

// SbTMVP CUs (MRG_TYPE_SUBPU_ATMVP): InterPrediction::xSubPuMC (:438-549) joins 8x8 sub-blocks with equal motion into runs along the
// longer CU side (split once more where a run longer than 16 is not a multiple of 16) and predicts each run like a CU of its own — the
// run geometry matters because MV clipping is relative to the run's position.  emit( const b200_pu& ) is called per run.
template<class WpIdxOf, class Emit>
inline FlattenPuResult flattenSbTmvp( const CodingUnit& cu, const SlotMap& sm, WpIdxOf wpIdxOf, Emit emit )
{
  const Position puPos = cu.lumaPos(); const Size puSize = cu.lumaSize();
  const int numPartLine = std::max<SizeType>( puSize.width >> ATMVP_SUB_BLOCK_SIZE, 1u ), numPartCol = std::max<SizeType>( puSize.height >> ATMVP_SUB_BLOCK_SIZE, 1u );
  const int puHeight = numPartCol == 1 ? puSize.height : 1 << ATMVP_SUB_BLOCK_SIZE, puWidth = numPartLine == 1 ? puSize.width : 1 << ATMVP_SUB_BLOCK_SIZE;
  CodingUnit sub;
  memset( (void*) &sub, 0, sizeof( sub ) );            // plain-data class; the reference's local subCu relies on the same fields it sets below
  sub.cs = cu.cs; sub.slice = cu.slice; sub.pps = cu.pps; sub.sps = cu.sps;
  sub.ctuData = cu.ctuData;
  sub.setChType( cu.chType() ); sub.setPredMode( cu.predMode() ); sub.UnitArea::operator=( cu );
  sub.setMergeType( MRG_TYPE_DEFAULT_N ); sub.setAffineFlag( false ); sub.setGeoFlag( false );
  sub.setBcwIdx( cu.BcwIdx() ); sub.setImv( cu.imv() ); sub.setSmvdMode( cu.smvdMode() );
  const bool verMC = puSize.height > puSize.width;
  const int fstStart = !verMC ? puPos.y : puPos.x, secStart = !verMC ? puPos.x : puPos.y;
  const int fstEnd = !verMC ? puPos.y + puSize.height : puPos.x + puSize.width, secEnd = !verMC ? puPos.x + puSize.width : puPos.y + puSize.height;
  const int fstStep = !verMC ? puHeight : puWidth, secStep = !verMC ? puWidth : puHeight;
  auto one = [&]( int x, int y, int dx, int dy ) -> FlattenPuResult
  {
    new ( &static_cast<UnitArea&>( sub ) ) UnitArea( cu.chromaFormat, Area( x, y, dx, dy ) );
    b200_pu r;
    const FlattenPuResult rc = flattenPU( sub, sm, wpIdxOf, r, true );
    if( rc == FLATTEN_PU_OK ) emit( r );
    return rc;
  };
  for( int fstDim = fstStart; fstDim < fstEnd; fstDim += fstStep )
    for( int secDim = secStart; secDim < secEnd; secDim += secStep )
    {
      int x = !verMC ? secDim : fstDim, y = !verMC ? fstDim : secDim;
      const MotionInfo& curMi = cu.getMotionInfo( Position{ x, y } );
      int length = secStep, later = secDim + secStep;
      while( later < secEnd )
      {
        const MotionInfo& laterMi = !verMC ? cu.getMotionInfo( Position{ later, fstDim } ) : cu.getMotionInfo( Position{ fstDim, later } );
        if( !( laterMi == curMi ) ) break;
        length += secStep; later += secStep;
      }
      int dx = !verMC ? length : puWidth, dy = !verMC ? puHeight : length;
      sub = curMi;
      if( !verMC && ( dx & 15 ) && dx > 16 )     { const int part = dx & ~15; if( FlattenPuResult rc = one( x, y, part, dy ) ) return rc; x += part; dx -= part; }
      else if( verMC && ( dy & 15 ) && dy > 16 ) { const int part = dy & ~15; if( FlattenPuResult rc = one( x, y, dx, part ) ) return rc; y += part; dy -= part; }
      if( FlattenPuResult rc = one( x, y, dx, dy ) ) return rc;
      secDim = later - secStep;
    }
  return FLATTEN_PU_OK;
}

}   // namespace b200glue
