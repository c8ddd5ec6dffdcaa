// ref_seam.h — the DecLibRecon seam, executed (SURVEY 8c level L1).  TEST / BENCH INFRASTRUCTURE ONLY; included at the end of ref_shim.cpp.
//
// A complete synthetic *parsed* Picture is built the way CABACReader / DecSlice leave one behind — through the reference's own allocators and
// helpers: Partitioner::initCtu / canSplit / splitCurrArea / nextPart (the coding_tree walk, CABACReader.cpp:477-654, with the split, mode and
// syntax decisions drawn from a seeded RNG instead of the arithmetic decoder), CodingStructure::addCU / addTU / addEmptyTUs, Partitioner::setCUData,
// quantised levels written in place into the reconstruction plane (CABACReader.cpp:2457-2478, TS: :2868), CtuData SAO / ALF parameters, slice /
// picture-header / SPS / PPS / APS state.  Motion is left as *syntax* (merge indices, MVDs, MVP indices): deriving it is the back end's MIDER stage.
// On that Picture run either
//   (i)  the reference's own DecLibRecon::create / decompressPicture / waitForPrevDecompressedPic (DecLibRecon.cpp:127,429,684) with a
//        ThreadPool of N threads — BASELINE.md level B1, the CPU arm of bench.py — or
//   (ii) b200glue::DecLibReconB200 with the same three calls (the GPU back end behind the same seam),
// and compare reconstruction planes and the collocated motion (colMotion, what later pictures' TMVP reads) bit for bit.
#pragma once
#include "DecoderLib/DecLibRecon.h"
#include "CommonLib/UnitPartitioner.h"

namespace vvdec { int signalModeCons( const CodingStructure& cs, const Slice* slice, const PartSplit split, const Partitioner& partitioner, const ModeType modeTypeParent ); }   // CABACReader.cpp:630

namespace seam
{
using namespace vvdec;
static const bool kTrace = getenv( "SEAM_TRACE" ) != nullptr;
} // namespace seam
#include <execinfo.h>
#include <signal.h>
namespace seam {
static void segvHandler( int sig ) { void* bt[64]; const int n = backtrace( bt, 64 ); fprintf( stderr, "signal %d, backtrace:\n", sig ); backtrace_symbols_fd( bt, n, 2 ); _exit( 139 ); }
static const bool kSegv = getenv( "SEAM_BACKTRACE" ) ? ( signal( SIGSEGV, segvHandler ), true ) : false;
#define SEAM_TR( ... ) do { if( seam::kTrace ) { fprintf( stderr, __VA_ARGS__ ); fflush( stderr ); } } while( 0 )

static const uint8_t kDiagX[16] = { 0,0,1,0,1,2,0,1,2,3,1,2,3,2,3,3 }, kDiagY[16] = { 0,1,0,2,1,0,3,2,1,0,3,2,1,3,2,3 };   // 4x4 diagonal scan (g_scanOrder[SCAN_GROUPED_4x4][2][2])

struct Gen
{
  FakePicture& fp; CodingStructure& cs; Slice* sl; const ref_seam_cfg& cfg; std::mt19937 rng;
  Partitioner part;
  int qp;
  Gen( FakePicture& f, Slice* s, const ref_seam_cfg& c ) : fp( f ), cs( *f.pic.cs ), sl( s ), cfg( c ), rng( c.seed ), qp( c.qp ) {}

  unsigned rnd( unsigned n ) { return n ? unsigned( rng() % n ) : 0; }
  bool pct( int p ) { return int( rng() % 100 ) < p; }
  bool tool( int t ) const { return ( cfg.tools & t ) != 0; }

  // ------------------------------------------------------------------ coding_tree_unit (CABACReader.cpp:128)
  void ctu( unsigned a )
  {
    const PreCalcValues& pcv = *cs.pcv;
    const unsigned x = a % pcv.widthInCtus, y = a / pcv.widthInCtus;
    const UnitArea ctuArea( pcv.chrFormat, Area( x * pcv.maxCUWidth, y * pcv.maxCUHeight, pcv.maxCUWidth, pcv.maxCUHeight ) );
    CtuData& cd = cs.getCtuData( a );
    cd.slice = sl; cd.pps = fp.pps.get(); cd.sps = fp.sps.get(); cd.ph = fp.ph.get();                      // DecSlice.cpp:149-153
    part.initCtu( ctuArea, CH_L, cs, *sl );
    part.treeType = TREE_D; part.modeType = MODE_TYPE_ALL;
    codingTree();
  }

  // split_cu_mode (:679) with the decision drawn at random among what Partitioner::canSplit allows
  PartSplit chooseSplit()
  {
    bool canNo, canQt, canBh, canBv, canTh, canTv;
    part.canSplit( cs, canNo, canQt, canBh, canBv, canTh, canTv );
    PartSplit cand[5]; int n = 0;
    auto add = [&]( bool can, PartSplit s )
    {
      if( !can ) return;
      // local dual trees (mode_constraint :657): only when the configuration asks for them
      if( !tool( SEAM_LOCAL_DUAL_TREE ) && canNo && signalModeCons( cs, sl, s, part, part.modeType ) != LDT_MODE_TYPE_INHERIT ) return;
      cand[n++] = s;
    };
    add( canQt, CU_QUAD_SPLIT ); add( canBh, CU_HORZ_SPLIT ); add( canBv, CU_VERT_SPLIT ); add( canTh, CU_TRIH_SPLIT ); add( canTv, CU_TRIV_SPLIT );
    if( !n ) return CU_DONT_SPLIT;
    if( canNo )
    {
      const Area& b = part.currArea().blocks[part.chType];
      const int big = std::max( b.width, b.height ) << ( isChroma( part.chType ) ? 1 : 0 );
      const int pSplit = big >= 128 ? 97 : big >= 64 ? cfg.splitPct : big >= 32 ? cfg.splitPct * 3 / 4 : big >= 16 ? cfg.splitPct / 2 : cfg.splitPct / 3;
      if( !pct( pSplit ) ) return CU_DONT_SPLIT;
    }
    if( cand[0] == CU_QUAD_SPLIT && n > 1 && pct( 50 ) ) return CU_QUAD_SPLIT;
    return cand[rnd( n )];
  }

  // coding_tree (:477)
  void codingTree()
  {
    UnitArea currArea = part.currArea();
    const ModeType modeTypeParent = part.modeType;
    const PartSplit split = chooseSplit();
    if( split != CU_DONT_SPLIT )
    {
      const int val = signalModeCons( cs, sl, split, part, part.modeType );                                  // mode_constraint (:657)
      part.modeType = val == LDT_MODE_TYPE_SIGNAL ? ( pct( 50 ) ? MODE_TYPE_INTRA : MODE_TYPE_INTER ) : val == LDT_MODE_TYPE_INFER ? MODE_TYPE_INTRA : part.modeType;
      const bool chromaNotSplit = modeTypeParent == MODE_TYPE_ALL && part.modeType == MODE_TYPE_INTRA;
      if( part.treeType == TREE_D ) part.treeType = chromaNotSplit ? TREE_L : TREE_D;
      part.splitCurrArea( split, cs );
      do
      {
        if( cs.area.blocks[part.chType].contains( part.currArea().blocks[part.chType].pos() ) ) codingTree();
      } while( part.nextPart( cs ) );
      part.exitCurrSplit( cs );
      if( chromaNotSplit )
      {
        part.chType = CHANNEL_TYPE_CHROMA; part.treeType = TREE_C;
        part.updateNeighbors( cs );
        codingTree();
        part.chType = CHANNEL_TYPE_LUMA; part.treeType = TREE_D;
      }
      part.modeType = modeTypeParent;
      return;
    }
    TreeType treeType = part.treeType;
    if( isChroma( part.chType ) )                                { currArea.Y() = CompArea(); treeType = TREE_C; }
    else if( part.isDualITree || part.treeType == TREE_L )       { currArea.Cb() = currArea.Cr() = CompArea(); treeType = TREE_L; }
    CodingUnit& cu = cs.addCU( currArea, part.chType, treeType, part.modeType, part.currPartLevel().cuLeft, part.currPartLevel().cuAbove );
    part.setCUData( cu );
    SEAM_TR( "cu %d,%d %dx%d ch%d tree%d mode%d\n", cu.blocks[cu.chType()].x, cu.blocks[cu.chType()].y, cu.blocks[cu.chType()].width, cu.blocks[cu.chType()].height, (int) cu.chType(), (int) cu.treeType(), (int) cu.modeType() );
    cu.slice = sl; cu.pps = fp.pps.get(); cu.sps = fp.sps.get(); cu.tileIdx = part.currTileIdx;
    if( isChroma( cu.chType() ) )
    {                                                                                                      // :590-603: chroma CU of a separate tree takes the co-located luma CU's QP
      const Position c( cu.chromaPos().offset( cu.chromaSize().width >> 1, cu.chromaSize().height >> 1 ) );
      const CodingUnit* col = cs.getCU( Position( c.x << 1, c.y << 1 ), CHANNEL_TYPE_LUMA );
      cu.qp = col ? col->qp : qp;
    }
    else
    {
      if( pct( 25 ) ) qp = std::min( 45, std::max( 20, qp + int( rnd( 9 ) ) - 4 ) );                     // cu_qp_delta at quantisation-group starts: a bounded random walk
      cu.qp = (int8_t) qp;
    }
    cu.chromaQpAdj = 0;
    codingUnit( cu );
    SEAM_TR( "   pred%d skip%d merge%d aff%d geo%d ciip%d mmvd%d dir%d imv%d bcw%d smvd%d | intra %d %d mrl%d mip%d bdpcm%d%d isp%d | rootCbf%d sbt%d lfnst%d mts%d qp%d\n", (int) cu.predMode(), cu.skip(), cu.mergeFlag(), cu.affineFlag(), cu.geoFlag(),
             cu.ciipFlag(), cu.mmvdFlag(), cu.interDir(), cu.imv(), cu.BcwIdx(), cu.smvdMode(), cu.intraDir[0], cu.intraDir[1], cu.multiRefIdx(), cu.mipFlag(), cu.bdpcmMode(), cu.bdpcmModeChroma(), cu.ispMode(),
             cu.rootCbf(), cu.sbtInfo(), cu.lfnstIdx(), cu.firstTU.mtsIdx( 0 ), cu.qp );
    if( isChromaEnabled( cs.pcv->chrFormat ) )                                                             // :620-634
      for( TransformUnit& tu : TUTraverser( &cu.firstTU, cu.lastTU->next ) )
      {
        if( tu.Cb().valid() ) { QpParam cQP( tu, COMPONENT_Cb, false ); tu.chromaQp[0] = cQP.Qp( false ); }
        if( tu.Cr().valid() ) { QpParam cQP( tu, COMPONENT_Cr, false ); tu.chromaQp[1] = cQP.Qp( false ); }
      }
  }

  // ------------------------------------------------------------------ coding_unit (:856)
  void codingUnit( CodingUnit& cu )
  {
    if( !sl->isIntra() )
    {
      const bool noSkip = ( cu.lwidth() == 4 && cu.lheight() == 4 ) || CU::isConsIntra( cu );             // cu_skip_flag (:908)
      if( cu.Y().valid() && !noSkip && pct( cfg.skipPct ) )
      {
        cu.setSkip( true ); cu.setColorTransform( false );
        cs.addEmptyTUs( part, cu );
        predictionUnit( cu );
        return;
      }
      // pred_mode (:1055)
      if( CU::isConsInter( cu ) ) cu.setPredMode( MODE_INTER );
      else if( ( cu.lwidth() == 4 && cu.lheight() == 4 ) || CU::isConsIntra( cu ) || !cu.Y().valid() ) cu.setPredMode( MODE_INTRA );
      else cu.setPredMode( pct( cfg.intraPct ) ? MODE_INTRA : MODE_INTER );
    }
    else cu.setPredMode( MODE_INTRA );
    cuPredData( cu );
    cuResidual( cu );
  }

  // ------------------------------------------------------------------ cu_pred_data (:1143), intra part
  void cuPredData( CodingUnit& cu )
  {
    if( !CU::isIntra( cu ) )
    {
      predictionUnit( cu );
      if( !cu.mergeFlag() )
      {
        // amvr_mode / affine_amvr_mode (:991,:1031): only signalled with a non-zero MVD
        if( cu.affineFlag() ) { if( cu.sps->getAffineAmvrEnabledFlag() && CU::hasSubCUNonZeroAffineMVd( cu ) && pct( 30 ) ) cu.setImv( 1 + rnd( 2 ) ); }
        else if( cu.sps->getAMVREnabledFlag() && CU::hasSubCUNonZeroMVd( cu ) && pct( 40 ) ) cu.setImv( 1 + rnd( 3 ) );      // 1 integer, 2 four-sample, 3 half-sample (IMV_HPEL)
        if( CU::isBcwIdxCoded( cu ) && pct( 30 ) ) cu.setBcwIdx( 1 + rnd( sl->getCheckLDC() ? 4 : 2 ) );                      // cu_bcw_flag (:1180), internal index
      }
      return;
    }
    if( isLuma( cu.chType() ) )
    {
      if( CU::bdpcmAllowed( cu, COMPONENT_Y ) && tool( SEAM_BDPCM ) && pct( 4 ) ) { cu.setBdpcmMode( 1 + rnd( 2 ) ); cu.intraDir[0] = cu.bdpcmMode() == 2 ? VER_IDX : HOR_IDX; }
      else if( cu.sps->getUseMIP() && cu.lwidth() <= 64 && cu.lheight() <= 64 && pct( 10 ) )
      {                                                                                                    // mip_flag / mip_pred_mode (:3125)
        cu.setMipFlag( true ); cu.setMipTransposedFlag( pct( 50 ) );
        cu.intraDir[0] = (int8_t) rnd( getNumModesMip( cu.Y() ) );
      }
      else
      {
        // extend_ref_line (:1242), isp_mode (:2541), MPM / remaining mode (:1270)
        const bool firstLineOfCtu = ( cu.ly() & cs.pcv->maxCUHeightMask ) == 0;
        if( cu.sps->getUseMRL() && !firstLineOfCtu && pct( 12 ) ) cu.setMultiRefIdx( 1 + rnd( 2 ) );
        if( !cu.multiRefIdx() && cu.sps->getUseISP() && pct( cfg.ispPct ) )
        {
          const int allowed = CU::canUseISPSplit( cu, COMPONENT_Y );                                        // 0 none, 1 / 2 one direction, 4 both (UnitTools.cpp:343)
          if( allowed ) cu.setIspMode( allowed > 2 ? 1 + rnd( 2 ) : allowed );
        }
        unsigned mpm[NUM_MOST_PROBABLE_MODES];
        PU::getIntraMPMs( cu, mpm );
        if( cu.multiRefIdx() ) cu.intraDir[0] = (int8_t) mpm[1 + rnd( 5 )];                                 // an MPM that is not planar
        else if( pct( 50 ) ) cu.intraDir[0] = (int8_t) mpm[rnd( 6 )];
        else cu.intraDir[0] = (int8_t) rnd( NUM_LUMA_MODE );
      }
    }
    if( ( isChroma( cu.chType() ) || !CU::isSepTree( cu ) ) && isChromaEnabled( cu.chromaFormat ) )
    {
      if( CU::bdpcmAllowed( cu, COMPONENT_Cb ) && tool( SEAM_BDPCM ) && pct( 3 ) ) { cu.setBdpcmModeChroma( 1 + rnd( 2 ) ); cu.intraDir[1] = cu.bdpcmModeChroma() == 2 ? VER_IDX : HOR_IDX; }
      else if( cu.sps->getUseLMChroma() && CU::checkCCLMAllowed( cu ) && pct( 25 ) )
      {                                                                                                    // intra_chroma_lmc_mode (:1384)
        int lm[10]; PU::getLMSymbolList( cu, lm );
        cu.intraDir[1] = (int8_t) lm[rnd( 3 )];
      }
      else if( pct( 40 ) ) cu.intraDir[1] = DM_CHROMA_IDX;
      else { unsigned cand[NUM_CHROMA_MODE]; PU::getIntraChromaCandModes( cu, cand ); cu.intraDir[1] = (int8_t) cand[rnd( 4 )]; }
    }
  }

  Mv randMvd( int sigmaQpel )
  {
    std::normal_distribution<float> nd( 0.f, float( sigmaQpel ) );
    return Mv( int( nd( rng ) ), int( nd( rng ) ) );
  }

  // ------------------------------------------------------------------ prediction_unit (:1568) + merge_data (:1732)
  void predictionUnit( CodingUnit& cu )
  {
    const SPS& sps = *cu.sps;
    if( cu.skip() ) cu.setMergeFlag( true ); else cu.setMergeFlag( pct( cfg.mergePct ) );
    if( cu.mergeFlag() )
    {
      // subblock_merge_flag (:1681)
      if( cs.picHeader->getMaxNumAffineMergeCand() > 0 && cu.lwidth() >= 8 && cu.lheight() >= 8 && pct( cfg.affinePct ) )
      {
        cu.setAffineFlag( true ); cu.setMergeIdx( rnd( cs.picHeader->getMaxNumAffineMergeCand() ) );
        return;
      }
      const bool ciipAvailable = sps.getUseCiip() && !cu.skip() && cu.lwidth() < 128 && cu.lheight() < 128 && cu.Y().area() >= 64;
      const bool geoAvailable  = sps.getUseGeo() && sl->isInterB() && cu.lwidth() >= GEO_MIN_CU_SIZE && cu.lheight() >= GEO_MIN_CU_SIZE && cu.lwidth() <= GEO_MAX_CU_SIZE
                                 && cu.lheight() <= GEO_MAX_CU_SIZE && cu.lwidth() < 8 * cu.lheight() && cu.lheight() < 8 * cu.lwidth();
      bool regular = true;
      if( geoAvailable || ciipAvailable ) regular = !pct( 25 );
      if( regular )
      {
        if( sps.getUseMMVD() && pct( 20 ) )
        {                                                                                                  // mmvd_merge_idx (:1882)
          cu.setMmvdFlag( true );
          const int base = sps.getMaxNumMergeCand() > 1 ? rnd( MMVD_BASE_MV_NUM ) : 0;
          cu.mmvdIdx = uint8_t( base * MMVD_MAX_REFINE_NUM + rnd( MMVD_REFINE_STEP ) * 4 + rnd( 4 ) );
          return;
        }
        cu.setMergeIdx( rnd( sps.getMaxNumMergeCand() ) );
        return;
      }
      if( geoAvailable && ciipAvailable ) cu.setCiipFlag( pct( 50 ) ); else if( ciipAvailable ) cu.setCiipFlag( true );
      if( cu.ciipFlag() ) { cu.intraDir[0] = PLANAR_IDX; cu.intraDir[1] = DM_CHROMA_IDX; cu.setMergeIdx( rnd( sps.getMaxNumMergeCand() ) ); return; }
      cu.setGeoFlag( true );                                                                                // merge_idx (:1808), geo branch
      cu.geoSplitDir = (uint8_t) rnd( GEO_NUM_PARTITION_MODE );
      const int nGeo = sps.getMaxNumGeoCand();
      const int c0 = rnd( nGeo ); int c1 = rnd( nGeo - 1 ); c1 += c1 >= c0 ? 1 : 0;
      cu.setGeoMergeIdx0( c0 ); cu.setGeoMergeIdx1( c1 );
      return;
    }
    // AMVP: inter_pred_idc (:1917), affine_flag (:1694), smvd_mode (:1662), ref_idx (:1948), mvd_coding, mvp_flag
    int dir = 1;
    if( sl->isInterB() )
    {
      const bool biOk = !PU::isBipredRestriction( cu );
      dir = ( biOk && pct( cfg.biPct ) ) ? 3 : 1 + rnd( 2 );
    }
    cu.setInterDir( dir );
    if( sps.getUseAffine() && cu.lwidth() >= 16 && cu.lheight() >= 16 && pct( cfg.affinePct ) )
    {
      cu.setAffineFlag( true );
      if( sps.getUseAffineType() ) cu.setAffineType( pct( 40 ) ? AFFINEMODEL_6PARAM : AFFINEMODEL_4PARAM );
    }
    if( dir == 3 && !cu.affineFlag() && sps.getUseSMVD() && !cs.picHeader->getMvdL1ZeroFlag() && sl->getBiDirPred() && pct( 15 ) ) cu.setSmvdMode( 1 );
    const int sig = cfg.mvdSigmaQpel;
    for( int l = 0; l < 2; l++ )
    {
      if( !( dir & ( 1 << l ) ) ) continue;
      if( l == 1 && cu.smvdMode() == 1 ) { cu.mvpIdx[1] = (uint8_t) rnd( 2 ); continue; }
      cu.refIdx[l] = cu.smvdMode() ? (int8_t) sl->getSymRefIdx( l ) : (int8_t) rnd( sl->getNumRefIdx( RefPicList( l ) ) );
      if( !( l == 1 && cs.picHeader->getMvdL1ZeroFlag() && dir == 3 ) )
      {
        cu.mv[l][0] = pct( 25 ) ? Mv() : randMvd( sig );
        if( cu.affineFlag() ) { cu.mv[l][1] = randMvd( std::max( 1, sig / 4 ) ); if( cu.affineType() == AFFINEMODEL_6PARAM ) cu.mv[l][2] = randMvd( std::max( 1, sig / 4 ) ); }
      }
      cu.mvpIdx[l] = (uint8_t) rnd( 2 );
    }
    if( cu.smvdMode() )
    {                                                                                                      // :1649-1657
      cu.mv[1][0].set( -cu.mv[0][0].hor, -cu.mv[0][0].ver );
      cu.refIdx[1] = (int8_t) sl->getSymRefIdx( 1 );
    }
  }

  // ------------------------------------------------------------------ cu_residual (:1404) / transform_tree (:2012) / transform_unit (:2148)
  bool wantLfnst = false, wantMts = false;
  bool anyTs = false, lfnstLast = false, mtsLast = false;

  void cuResidual( CodingUnit& cu )
  {
    if( !CU::isIntra( cu ) )
    {
      cu.setRootCbf( cu.mergeFlag() ? true : pct( cfg.rootCbfPct ) );                                      // rqt_root_cbf (:1459)
      if( cu.rootCbf() )
      {                                                                                                    // sbt_mode (:1476)
        const uint8_t allowed = CU::checkAllowedSbt( cu );
        if( allowed && pct( 20 ) )
        {
          uint8_t opts[4]; int n = 0;
          for( uint8_t k = SBT_VER_HALF; k <= SBT_HOR_QUAD; k++ ) if( CU::targetSbtAllowed( k, allowed ) ) opts[n++] = k;
          CU::setSbtIdx( cu, opts[rnd( n )] ); CU::setSbtPos( cu, pct( 50 ) ? SBT_POS1 : SBT_POS0 );
        }
      }
      if( !cu.rootCbf() ) { cs.addEmptyTUs( part, cu ); return; }
    }
    else cu.setRootCbf( true );
    // the CU-level transform tools are signalled after the TUs and constrained by their coefficient positions (residual_lfnst_mode :2578,
    // mts_idx :2508): decide first, then draw coefficients that satisfy the constraints
    const bool isIntra = CU::isIntra( cu );
    const Size lSize = cu.blocks[CU::isSepTree( cu ) && isChroma( cu.chType() ) ? 1 : 0].lumaSize( cu.chromaFormat );
    wantLfnst = isIntra && cu.sps->getUseLFNST() && pct( 25 )
                && !( cu.ispMode() && !CU::canUseLfnstWithISP( cu, cu.chType() ) ) && !( cu.mipFlag() && !allowLfnstWithMip( cu.lumaSize() ) )
                && !( isChroma( cu.chType() ) && std::min( cu.blocks[1].width, cu.blocks[1].height ) < 4 )
                && lSize.width <= cu.sps->getMaxTbSize() && lSize.height <= cu.sps->getMaxTbSize();
    wantMts = !wantLfnst && CU::isMTSAllowed( cu, COMPONENT_Y ) && pct( 25 );
    anyTs = false; lfnstLast = false; mtsLast = false;
    transformTree( cu );
    if( wantLfnst && ( lfnstLast || cu.ispMode() ) && !anyTs ) cu.setLfnstIdx( 1 + rnd( 2 ) );
    if( wantMts && mtsLast && cu.firstTU.mtsIdx( COMPONENT_Y ) != MTS_SKIP && TU::getCbf( cu.firstTU, COMPONENT_Y ) ) cu.firstTU.setMtsIdx( COMPONENT_Y, MTS_DST7_DST7 + rnd( 4 ) );
    bool rootCbf = false;
    for( const auto& blk : cu.blocks ) if( blk.valid() ) rootCbf |= cu.planeCbf( blk.compID() );
    cu.setRootCbf( rootCbf );
  }

  void transformTree( CodingUnit& cu )
  {
    const UnitArea& area = part.currArea();
    bool split = area.Y().width > part.maxTrSize || area.Y().height > part.maxTrSize;
    const PartSplit ispType = CU::getISPType( cu, getFirstComponentOfChannel( part.chType ) );
    split |= ( cu.sbtInfo() || ispType != TU_NO_ISP ) && part.currTrDepth == 0;
    if( split )
    {
      if( ispType == TU_NO_ISP && !cu.sbtInfo() ) part.splitCurrArea( TU_MAX_TR_SPLIT, cs );
      else if( ispType != TU_NO_ISP )            part.splitCurrArea( ispType, cs );
      else                                       part.splitCurrArea( PartSplit( CU::getSbtTuSplit( cu ) ), cs );
      do { transformTree( cu ); } while( part.nextPart( cs ) );
      part.exitCurrSplit( cs );
      return;
    }
    TransformUnit& tu = cs.addTU( getArea( *sl, area, part.chType, part.treeType ), part.chType, cu );
    transformUnit( tu );
  }

  void transformUnit( TransformUnit& tu )
  {
    const UnitArea& area = part.currArea();
    const unsigned trDepth = part.currTrDepth;
    CodingUnit& cu = *tu.cu;
    bool cbfCb = false, cbfCr = false;
    const bool chromaCbfISP = isChromaEnabled( area.chromaFormat ) && area.blocks[COMPONENT_Cb].valid() && cu.ispMode();
    const bool tuNoResidual = TU::checkTuNoResidual( tu, part.currPartIdx() );
    if( area.chromaFormat != CHROMA_400 && area.blocks[COMPONENT_Cb].valid() && ( !CU::isSepTree( cu ) || part.chType == CHANNEL_TYPE_CHROMA ) && ( !cu.ispMode() || chromaCbfISP ) )
      if( !( cu.sbtInfo() && tuNoResidual ) ) { cbfCb = pct( cfg.cbfPct ); cbfCr = pct( cfg.cbfPct ); }
    const bool sigChroma = cbfCb || cbfCr;
    if( !isChroma( part.chType ) )
    {
      bool cbfY;
      if( !CU::isIntra( cu ) && trDepth == 0 && !sigChroma ) cbfY = true;
      else if( cu.sbtInfo() && tuNoResidual ) cbfY = false;
      else if( cu.sbtInfo() && !sigChroma ) cbfY = true;
      else if( cu.ispMode() )
      {
        const int nTus = cu.ispMode() == HOR_INTRA_SUBPARTITIONS ? cu.lheight() >> getLog2( tu.lheight() ) : cu.lwidth() >> getLog2( tu.lwidth() );
        bool rootSoFar = false;
        if( (int) part.currPartIdx() == nTus - 1 ) for( const TransformUnit& t : cTUTraverser( &cu.firstTU, cu.lastTU ) ) rootSoFar |= TU::getCbf( t, COMPONENT_Y );
        cbfY = ( (int) part.currPartIdx() == nTus - 1 && !rootSoFar ) ? true : pct( std::max( cfg.cbfPct, 50 ) );
      }
      else cbfY = pct( std::max( cfg.cbfPct, 60 ) );
      TU::setCbf( tu, COMPONENT_Y, cbfY );
    }
    if( area.chromaFormat != CHROMA_400 && ( !cu.ispMode() || chromaCbfISP ) ) { TU::setCbf( tu, COMPONENT_Cb, cbfCb ); TU::setCbf( tu, COMPONENT_Cr, cbfCr ); }
    cu.setPlaneCbf( COMPONENT_Y,  cu.planeCbf( COMPONENT_Y )  || TU::getCbf( tu, COMPONENT_Y ) );
    cu.setPlaneCbf( COMPONENT_Cb, cu.planeCbf( COMPONENT_Cb ) || TU::getCbf( tu, COMPONENT_Cb ) );
    cu.setPlaneCbf( COMPONENT_Cr, cu.planeCbf( COMPONENT_Cr ) || TU::getCbf( tu, COMPONENT_Cr ) );
    const bool lumaOnly = cu.chromaFormat == CHROMA_400 || !tu.blocks[COMPONENT_Cb].valid();
    const bool cbfLuma = TU::getCbf( tu, COMPONENT_Y ), cbfChroma = lumaOnly ? false : sigChroma;
    if( !( cbfLuma || cbfChroma ) ) return;
    if( !lumaOnly && cu.sps->getJointCbCrEnabledFlag() )
    {                                                                                                      // joint_cb_cr (:2349)
      const int mask = ( cbfCb ? 2 : 0 ) + ( cbfCr ? 1 : 0 );
      if( ( ( CU::isIntra( cu ) && mask ) || mask == 3 ) && pct( 25 ) ) { tu.jointCbCr = mask; cu.setPlaneCbf( COMPONENT_Cb, true ); cu.setPlaneCbf( COMPONENT_Cr, true ); }
    }
    if( cbfLuma ) residual( tu, COMPONENT_Y );
    if( !lumaOnly ) for( int c = 1; c <= 2; c++ ) if( TU::getCbf( tu, ComponentID( c ) ) ) residual( tu, ComponentID( c ) );
  }

  int level( bool big )
  {
    std::geometric_distribution<int> gd( big ? 0.08 : 0.35 );
    int v = 1 + gd( rng );
    if( pct( 2 ) ) v += rnd( 1500 );
    return pct( 50 ) ? -v : v;
  }

  // residual_coding (:2362) / residual_codingTS (:2863): what they leave in the plane and in maxScanPos
  void residual( TransformUnit& tu, const ComponentID compID )
  {
    CodingUnit& cu = *tu.cu;
    if( compID == COMPONENT_Cr && tu.jointCbCr == 3 ) return;
    const CompArea& blk = tu.blocks[compID];
    const int w = blk.width, h = blk.height;
    const bool bdpcm = isLuma( compID ) ? cu.bdpcmMode() != 0 : cu.bdpcmModeChroma() != 0;
    bool ts = bdpcm;
    if( !wantLfnst && TU::isTSAllowed( tu, compID ) && tool( SEAM_TS ) && pct( 6 ) ) ts = true;              // ts_flag (:2493)
    tu.setMtsIdx( compID, ts ? MTS_SKIP : MTS_DCT2_DCT2 );
    PelBuf full = cs.getRecoBuf( blk );
    if( ts )
    {
      anyTs = true;
      full.memset( 0 );
      int maxX = 0, maxY = 0; const int n = 1 + rnd( std::max( 1, w * h / 3 ) );
      for( int i = 0; i < n; i++ ) { const int x = rnd( w ), y = rnd( h ); full.at( x, y ) = (Pel) std::max( -400, std::min( 400, level( true ) ) ); maxX = std::max( maxX, x ); maxY = std::max( maxY, y ); }
      tu.maxScanPosX[compID] = bdpcm ? w : maxX; tu.maxScanPosY[compID] = bdpcm ? h : maxY;
      return;
    }
    const int cgW = std::min( 4, w ) == 4 && std::min( 4, h ) == 4 ? 4 : ( 1 << g_log2SbbSize[getLog2( w )][getLog2( h )][0] );
    const int cgH = cgW == 4 && h >= 4 && w >= 4 ? 4 : ( 1 << g_log2SbbSize[getLog2( w )][getLog2( h )][1] );
    // zero-out regions: 32 for 64-sized DCT-2 (JVET_C0024_ZERO_OUT_TH); 16 for MTS and for SBT's implicit DST-7 / DCT-8 at 32 (skipBlkPreCond :2414)
    int limW = std::min( w, 32 ), limH = std::min( h, 32 );
    const bool sbtMts = isLuma( compID ) && cu.sps->getUseMTS() && cu.sbtInfo() != 0 && w <= 32 && h <= 32;
    if( ( isLuma( compID ) && wantMts ) || sbtMts ) { limW = std::min( limW, 16 ); limH = std::min( limH, 16 ); }
    if( cu.ispMode() && isLuma( compID ) && cu.sps->getUseMTS() && cu.sps->getUseIntraMTS() == false ) {}   // implicit MTS keeps the DCT-2 zero-out of 32
    const bool lfnstBlk = wantLfnst && w >= 4 && h >= 4;
    int nCgX = std::max( 1, limW / cgW ), nCgY = std::max( 1, limH / cgH );
    int mx = 0, my = 0;
    if( !lfnstBlk ) { std::geometric_distribution<int> gd( 0.45 ); mx = std::min( nCgX - 1, gd( rng ) ); my = std::min( nCgY - 1, gd( rng ) ); }
    const bool dcOnly = !lfnstBlk && mx == 0 && my == 0 && pct( 35 );
    const int maxX = ( mx + 1 ) * cgW, maxY = ( my + 1 ) * cgH;
    PelBuf pb = cs.getRecoBuf( CompArea( compID, blk.pos(), Size( maxX, maxY ) ) );
    pb.memset( 0 );
    if( dcOnly ) { pb.at( 0, 0 ) = (Pel) level( true ); tu.maxScanPosX[compID] = 0; tu.maxScanPosY[compID] = 0; return; }
    bool nonDc = false;
    if( lfnstBlk )
    {
      // all significant coefficients inside the first 8 (4x4 / 8x8 blocks) or 16 scan positions (violatesLfnstConstrained :2392)
      const int maxPos = ( ( w == 4 && h == 4 ) || ( w == 8 && h == 8 ) ) ? 8 : 16;
      const int n = 1 + rnd( maxPos );
      for( int i = 0; i < n; i++ ) { const int k = rnd( maxPos ); pb.at( kDiagX[k], kDiagY[k] ) = (Pel) level( true ); nonDc |= k > 0; }
      if( !nonDc ) { const int k = 1 + rnd( maxPos - 1 ); pb.at( kDiagX[k], kDiagY[k] ) = (Pel) level( false ); nonDc = true; }
    }
    else
    {
      // a coefficient in the coefficient group that defines maxScanPos, the rest sparse with a low-frequency bias
      const int cx = mx * cgW + rnd( cgW ), cy = my * cgH + rnd( cgH );
      pb.at( ( cx | cy ) ? cx : std::min( 1, maxX - 1 ), ( cx | cy ) ? cy : ( maxX > 1 ? 0 : std::min( 1, maxY - 1 ) ) ) = (Pel) level( false );
      const int n = rnd( std::max( 2, maxX * maxY / 4 ) );
      for( int i = 0; i < n; i++ )
      {
        const int x = std::min( maxX - 1, int( rnd( maxX ) * rnd( 100 ) / 100 ) ), y = std::min( maxY - 1, int( rnd( maxY ) * rnd( 100 ) / 100 ) );
        pb.at( x, y ) = (Pel) level( x + y < 3 );
      }
      nonDc = true;
      if( maxX == 1 && maxY == 1 ) nonDc = false;
    }
    tu.maxScanPosX[compID] = maxX - 1; tu.maxScanPosY[compID] = maxY - 1;
    if( w >= 4 && h >= 4 ) lfnstLast |= nonDc;                                                              // lfnstLastScanPos (:2395-2398), LFNST_LAST_SIG_LUMA / CHROMA = 1
    if( isLuma( compID ) ) mtsLast |= nonDc;                                                                // mtsLastScanPos (:2400-2403)
  }
};

// ---------------------------------------------------------------------------------------------------------------- picture state
// the four reference pictures, shared by every handle built from the same planes (a decoder's references stay in the DPB — and on the device —
// from picture to picture; rebuilding them per handle would charge four 25 MB uploads to every picture of the GPU arm)
struct RefSet { b200_geom g; const int16_t* key[12]; uint64_t sum; std::unique_ptr<FakePicture> ref[4]; };
static std::shared_ptr<RefSet> refSetFor( const b200_geom& g, const int16_t* const* refs )
{
  static std::mutex m; static std::shared_ptr<RefSet> last;
  std::lock_guard<std::mutex> l( m );
  uint64_t sum = 0; for( int i = 0; i < 64; i++ ) sum = sum * 1315423911u + (uint16_t) refs[0][(size_t) i * 97 % ( (size_t) g.width * 8 )];
  if( last && !memcmp( &last->g, &g, sizeof( g ) ) && !memcmp( last->key, refs, sizeof( last->key ) ) && last->sum == sum ) return last;
  std::shared_ptr<RefSet> R( new RefSet ); R->g = g; memcpy( R->key, refs, sizeof( R->key ) ); R->sum = sum;
  const int pocs[4] = { 4, 0, 12, 16 };
  for( int s = 0; s < 4; s++ )
  {
    R->ref[s].reset( new FakePicture( g, 1 ) );
    int16_t* p3[3] = { (int16_t*) refs[s * 3], (int16_t*) refs[s * 3 + 1], (int16_t*) refs[s * 3 + 2] };
    R->ref[s]->setPlanes( g, p3 );
    Picture& rp = R->ref[s]->pic;
    rp.poc = pocs[s]; rp.progress = Picture::reconstructed; rp.reconDone.unlock(); rp.parseDone.unlock(); rp.stillReferenced = true;
    rp.extendPicBorder(); rp.borderExtStarted = true;
  }
  last = R;
  return R;
}

struct Pic
{
  b200_geom g;
  std::unique_ptr<FakePicture> cur; std::shared_ptr<RefSet> refs; FakePicture* ref[4] = { nullptr, nullptr, nullptr, nullptr };
  std::shared_ptr<APS> lmcsAps, slAps;
  std::shared_ptr<APS> alfAps[ALF_CTB_MAX_NUM_APS];
  ref_seam_cfg cfg;
  std::vector<MotionInfo> colMotionScratch;
};

static void setupSps( FakePicture& fp, const ref_seam_cfg& c, const b200_geom& g )
{
  SPS& sps = *fp.sps; PPS& pps = *fp.pps; PicHeader& ph = *fp.ph;
  sps.setUseBIO( c.tools & SEAM_BDOF ); sps.setUseDMVR( c.tools & SEAM_DMVR ); sps.setUseBcw( c.tools & SEAM_BCW );
  sps.setUseAffine( c.affinePct > 0 ); sps.setUseAffineType( true ); sps.setUsePROF( c.tools & SEAM_PROF );
  sps.setUseMMVD( c.tools & SEAM_MMVD ); sps.setUseGeo( c.tools & SEAM_GEO ); sps.setMaxNumGeoCand( 5 ); sps.setUseCiip( c.tools & SEAM_CIIP );
  sps.setUseSMVD( c.tools & SEAM_SMVD ); sps.setAMVREnabledFlag( c.tools & SEAM_AMVR ); sps.setAffineAmvrEnabledFlag( c.tools & SEAM_AMVR );
  sps.setMaxNumMergeCand( 6 ); sps.setSBTMVPEnabledFlag( false ); sps.setSPSTemporalMVPEnabledFlag( false );
  sps.setUseMTS( c.tools & SEAM_MTS ); sps.setUseIntraMTS( c.tools & SEAM_MTS ); sps.setUseInterMTS( c.tools & SEAM_MTS );
  sps.setUseLFNST( c.tools & SEAM_LFNST ); sps.setUseSBT( c.tools & SEAM_SBT );
  sps.setUseMRL( c.tools & SEAM_MRL ); sps.setUseMIP( c.tools & SEAM_MIP ); sps.setUseLMChroma( c.tools & SEAM_CCLM ); sps.setUseISP( c.ispPct > 0 );
  sps.setVerCollocatedChromaFlag( c.seed & 1 );
  sps.setJointCbCrEnabledFlag( c.tools & SEAM_JCCR );
  sps.setTransformSkipEnabledFlag( c.tools & SEAM_TS ); sps.setLog2MaxTransformSkipBlockSize( 5 ); sps.setBDPCMEnabledFlag( c.tools & SEAM_BDPCM );
  sps.setLog2MaxTbSize( 6 );
  sps.setUseSAO( c.tools & SEAM_SAO ); sps.setUseALF( c.tools & SEAM_ALF ); sps.setUseCCALF( c.tools & SEAM_ALF );
  sps.setUseReshaper( c.tools & SEAM_LMCS );
  sps.setUseDualITree( false ); sps.setIBCFlag( false ); sps.setUseColorTrans( false ); sps.setUseWrapAround( false );
  sps.setDepQuantEnabledFlag( true );
  sps.setScalingListFlag( c.tools & SEAM_SCALING_LIST ); sps.setDisableScalingMatrixForLfnstBlks( ( c.seed >> 2 ) & 1 );
  // partitioning limits (Partitioner::initCtu reads them, UnitPartitioner.cpp:158-188): index 0 intra slices, 1 inter slices, 2 chroma of a dual tree
  sps.setMinQTSizes( PartitionConstraints{ 8, 8, 4 } ); sps.setMaxMTTHierarchyDepths( PartitionConstraints{ 3, 3, 3 } );
  sps.setMaxBTSizes( PartitionConstraints{ 64, 128, 64 } ); sps.setMaxTTSizes( PartitionConstraints{ 64, 64, 32 } );
  sps.setInternalMinusInputBitDepth( 0 );
  SEAM_TR( "sps: flags\n" );
  {
    ChromaQpMappingTableParams p; p.m_qpBdOffset = sps.getQpBDOffset();
    sps.setChromaQpMappingTableFromParams( p ); sps.deriveChromaQPMappingTables();
  }
  // one tile, one slice: the tile maps the Partitioner / filters read (the no-partition branch of PPS::finalizePPSPartitioning, Slice.cpp:1579-1591)
  pps.setNoPicPartitionFlag( true ); pps.resetTileSliceInfo(); pps.setLog2CtuSize( getLog2( g.ctuSize ) );
  pps.setNumExpTileColumns( 1 ); pps.setNumExpTileRows( 1 ); pps.addTileColumnWidth( pps.getPicWidthInCtu() ); pps.addTileRowHeight( pps.getPicHeightInCtu() );
  pps.initTiles();
  SEAM_TR( "sps: tiles\n" );
  pps.setLoopFilterAcrossSlicesEnabledFlag( !( c.tools & SEAM_NO_LF_ACROSS_SLICES ) ); pps.setLoopFilterAcrossTilesEnabledFlag( true );
  pps.setUseWP( c.tools & SEAM_WP ); pps.setWPBiPred( c.tools & SEAM_WP );
  pps.setQpOffset( COMPONENT_Cb, 1 ); pps.setQpOffset( COMPONENT_Cr, -1 ); pps.setQpOffset( JOINT_CbCr, 0 );
  ph.setMaxNumAffineMergeCand( c.affinePct > 0 ? 5 : 0 ); ph.setEnableTMVPFlag( false ); ph.setMvdL1ZeroFlag( false );
  ph.setDisBdofFlag( false ); ph.setDisDmvrFlag( false ); ph.setDisProfFlag( false ); ph.setJointCbCrSignFlag( ( c.seed >> 1 ) & 1 );
  ph.setSplitConsOverrideFlag( false ); ph.setDisFracMMVD( false );
  ph.setVirtualBoundariesPresentFlag( c.tools & SEAM_VIRTUAL_BOUNDARIES ); ph.setNumVerVirtualBoundaries( 0 ); ph.setNumHorVirtualBoundaries( 0 );
  (void) g;
}

// prev: the picture (POC 8) the new one (then POC 10) predicts from as its first list-0 reference, in place of the shared reference picture of POC 4 — a chain of
// pictures as a decoder sees them: the reference may still be in reconstruction when this picture is handed to a recon instance
static Pic* build( const b200_geom* g, const ref_seam_cfg* c, const int16_t* const* refs, const b200_picture* filt, Pic* prev = nullptr )
{
  std::unique_ptr<Pic> P( new Pic );
  P->g = *g; P->cfg = *c;
  const int nCtus = ( ( g->width + g->ctuSize - 1 ) / g->ctuSize ) * ( ( g->height + g->ctuSize - 1 ) / g->ctuSize );
  const int nSl = std::max( 1, std::min( std::min( (int) c->numSlices, 16 ), nCtus ) );      // every slice holds at least one CTU
  P->cur.reset( new FakePicture( *g, nSl ) );
  FakePicture& cur = *P->cur;
  SEAM_TR( "build: picture created\n" );
  setupSps( cur, *c, *g );
  SEAM_TR( "build: sps done\n" );
  const bool intraPic = c->sliceType == 2;
  const int curPoc = prev ? 10 : 8;
  const int pocs[4] = { prev ? 8 : 4, 0, 12, 16 };
  if( !intraPic )
  {
    P->refs = refSetFor( *g, refs );
    for( int s = 0; s < 4; s++ ) P->ref[s] = P->refs->ref[s].get();
    if( prev ) P->ref[0] = prev->cur.get();
  }
  SEAM_TR( "build: refs done\n" );
  CodingStructure& cs = *cur.pic.cs;
  const PreCalcValues& pcv = *cs.pcv;
  auto sliceOfCtu = [&]( unsigned a ) { return (int) ( (uint64_t) a * nSl / pcv.sizeInCtus ); };
  const bool doLf = filt && ( filt->flags & B200_PIC_DEBLOCK ), doSao = filt && ( filt->flags & B200_PIC_SAO ) && ( c->tools & SEAM_SAO ), doAlf = filt && ( filt->flags & B200_PIC_ALF ) && ( c->tools & SEAM_ALF );
  const APS* apss[ALF_CTB_MAX_NUM_APS] = { nullptr };
  if( doAlf ) for( int i = 0; i < ALF_CTB_MAX_NUM_APS; i++ ) { P->alfAps[i] = std::make_shared<APS>(); P->alfAps[i]->setAPSId( i ); apss[i] = P->alfAps[i].get(); }
  for( int s = 0; s < nSl; s++ )
  {
  Slice* sl = cur.pic.slices[s];
  const bool odd = s & 1;
  sl->setSliceType( intraPic ? I_SLICE : c->sliceType == 1 ? P_SLICE : B_SLICE ); sl->setPOC( curPoc ); cur.pic.poc = curPoc;
  sl->setSliceQp( c->qp ); sl->setDepQuantEnabledFlag( c->tools & SEAM_DEPQUANT ); sl->setSignDataHidingEnabledFlag( false ); sl->setTSResidualCodingDisabledFlag( false );
  sl->setExplicitScalingListUsed( ( c->tools & SEAM_SCALING_LIST ) != 0 ); sl->setIndependentSliceIdx( s ); sl->setCheckLDC( false );
  if( ( c->tools & SEAM_SCALING_LIST ) && s == 0 )
  {
    // scaling_list_data: 28 matrices (2x2, 4x4, 8x8 coded sizes) + DC values for the lists of 16x16 and larger (Quant::setScalingListDec up-samples them)
    P->slAps = std::make_shared<APS>(); P->slAps->setAPSId( 1 ); P->slAps->setAPSType( SCALING_LIST_APS );
    std::mt19937 qr( c->seed * 131u + 7u );
    ScalingList& L = P->slAps->getScalingList();
    for( int id = 0; id < 28; id++ )
    {
      for( int& v : L.getScalingListVec( id ) ) v = 1 + (int) ( qr() % 255 );
      L.setScalingListDC( id, 1 + qr() % 255 );
    }
    cur.ph->setExplicitScalingListEnabledFlag( true ); cur.ph->setScalingListAPS( P->slAps );
  }
  if( !intraPic )
  {
    const int nL1 = sl->isInterB() ? 2 : 0;
    for( int l = 0; l < 2; l++ ) for( int i = 0; i < 2; i++ )
    { const int k = l * 2 + ( odd ? 1 - i : i ); sl->m_apcRefPicList[l][i] = &P->ref[k]->pic; sl->m_aiRefPOCList[l][i] = pocs[k]; sl->m_bIsUsedAsLongTerm[l][i] = false; }
    sl->setNumRefIdx( REF_PIC_LIST_0, 2 ); sl->setNumRefIdx( REF_PIC_LIST_1, nL1 );
    sl->resetWpScaling();
    if( c->tools & SEAM_WP )
    {
      // pred_weight_table (HLSyntaxReader.cpp: parsePredWeightTable): one luma and one chroma denominator per slice, weights / offsets per reference and component
      std::mt19937 wr( c->seed * 977u + 31u * s + 5u );
      const unsigned dL = wr() % 8, dC = wr() % 8;
      static thread_local WPScalingParam tab[NUM_REF_PIC_LIST_01][MAX_NUM_REF][MAX_NUM_COMPONENT];
      for( auto& a : tab ) for( auto& b : a ) for( auto& w : b ) w = WPScalingParam();
      for( int l = 0; l < 2; l++ ) for( int i = 0; i < 2; i++ ) for( int k = 0; k < 3; k++ )
      {
        WPScalingParam& w = tab[l][i][k];
        w.uiLog2WeightDenom = k ? dC : dL;
        w.bPresentFlag = wr() % 100 < 65;
        if( w.bPresentFlag ) { w.iWeight = ( 1 << w.uiLog2WeightDenom ) + (int) ( wr() % 65 ) - 32; w.iOffset = (int) ( wr() % ( k ? 21 : 41 ) ) - ( k ? 10 : 20 ); }
      }
      sl->setWpScaling( tab ); sl->initWpScaling( cur.sps.get() );
    }
    if( sl->isInterB() ) sl->setBiDirPred( true, odd ? 1 : 0, odd ? 1 : 0 );        // POC 4 and 12: the closest pair around POC 8 (Slice::setSMVDParam)
  }
  { SliceMap sm; for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) if( sliceOfCtu( a ) == s ) sm.addCtusToSlice( a % pcv.widthInCtus, a % pcv.widthInCtus + 1, a / pcv.widthInCtus, a / pcv.widthInCtus + 1, pcv.widthInCtus ); sl->setSliceMap( sm ); }
  cur.pic.stillReferenced = true; cur.pic.neededForOutput = true;
  SEAM_TR( "build: slice done\n" );
  // deblocking / SAO / ALF / LMCS state from the flat description (the same content the other tests use)
  sl->setDeblockingFilterDisable( !doLf );
  if( doLf )
  {
    auto off = [&]( int v, int k ) { return std::max( -12, std::min( 12, v + ( s ? ( s * 3 + k ) % 5 - 2 : 0 ) ) ); };      // other slices: other offsets
    sl->setDeblockingFilterBetaOffsetDiv2( off( filt->lfSlices[0].betaOffsetDiv2[0], 0 ) ); sl->setDeblockingFilterTcOffsetDiv2( off( filt->lfSlices[0].tcOffsetDiv2[0], 1 ) );
    sl->setDeblockingFilterCbBetaOffsetDiv2( off( filt->lfSlices[0].betaOffsetDiv2[1], 2 ) ); sl->setDeblockingFilterCbTcOffsetDiv2( off( filt->lfSlices[0].tcOffsetDiv2[1], 3 ) );
    sl->setDeblockingFilterCrBetaOffsetDiv2( off( filt->lfSlices[0].betaOffsetDiv2[2], 4 ) ); sl->setDeblockingFilterCrTcOffsetDiv2( off( filt->lfSlices[0].tcOffsetDiv2[2], 5 ) );
  }
  if( s == 0 ) for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
  {
    SAOBlkParam& bp = cs.getCtuData( a ).saoParam; bp.reset();
    if( !doSao ) continue;
    for( int k = 0; k < 3; k++ )
    {
      const b200_sao_ctu& sc = filt->sao[a];
      if( sc.type[k] == B200_SAO_OFF ) continue;
      bp[k].modeIdc = SAO_MODE_NEW; bp[k].typeIdc = sc.type[k]; bp[k].typeAuxInfo = sc.band[k];
      if( sc.type[k] == B200_SAO_BO ) for( int i = 0; i < 4; i++ ) bp[k].offset[( sc.band[k] + i ) & 31] = sc.offset[k][i];
      else for( int i = 0; i < 5; i++ ) bp[k].offset[i] = sc.offset[k][i];
    }
  }
  if( doAlf )
  {
    const b200_alf_tables* T = filt->alfTabs;
    const int nAps = T->numLumaSets - 16;
    const bool noCc = s % 3 == 2;                                    // every third slice: CC-ALF and the Cr filter off
    AlfApsIdVec ids;
    for( int i = 0; i < nAps; i++ )
    {
      if( s == 0 )
      {
        AlfSliceParam& p = P->alfAps[i]->getAlfAPSParam();
        memcpy( p.lumaCoeffFinal, T->lumaCoeff + (size_t) ( 16 + i ) * 1300, 2600 ); memcpy( p.lumaClippFinal, T->lumaClip + (size_t) ( 16 + i ) * 1300, 2600 );
        p.lumaFinalDone = true;
      }
      ids.push_back( odd ? nAps - 1 - i : i );                       // odd slices list the luma APSs backwards: the same CTU index means another filter set
    }
    sl->setNumAlfAps( nAps ); sl->setAlfApsIdsLuma( ids );
    if( s == 0 )
    {
      AlfSliceParam& pc = P->alfAps[7]->getAlfAPSParam();
      pc.numAlternativesChroma = T->numChromaAlts;
      memcpy( pc.chromaCoeff, T->chromaCoeff, T->numChromaAlts * 14 ); memcpy( pc.chrmClippFinal, T->chromaClip, T->numChromaAlts * 14 ); pc.chrmFinalDone = true;
      for( int k = 0; k < 2; k++ )
      {
        CcAlfFilterParam& cp = P->alfAps[6 - k]->getCcAlfAPSParam();
        for( int f = 0; f < T->numCc[k]; f++ ) memcpy( cp.ccAlfCoeff[k][f], T->ccCoeff[k] + f * 7, 14 );
        cp.ccAlfFilterCount[k] = (uint8_t) T->numCc[k];
      }
    }
    sl->setAlfApsIdChroma( 7 );
    sl->setCcAlfCbEnabledFlag( T->numCc[0] > 0 && !noCc ); sl->setCcAlfCrEnabledFlag( T->numCc[1] > 0 && !noCc ); sl->setCcAlfCbApsId( 6 ); sl->setCcAlfCrApsId( 5 );
    sl->setAlfApss( apss );
    for( int k = 0; k < 3; k++ ) sl->setAlfEnabledFlag( ComponentID( k ), !( noCc && k == 2 ) );
    for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
    {
      if( sliceOfCtu( a ) != s ) continue;
      CtuAlfData& d = cs.getCtuData( a ).alfParam;
      for( int k = 0; k < 3; k++ ) d.alfCtuEnableFlag[k] = ( filt->alf[a].enable[k] & 1 ) && !( noCc && k == 2 );     // what the parser leaves when the slice flag is off
      d.alfCtbFilterIndex = filt->alf[a].lumaSet;
      for( int k = 0; k < 2; k++ ) { d.alfCtuAlternative[k] = filt->alf[a].chromaAlt[k]; d.ccAlfFilterControl[k] = noCc ? 0 : filt->alf[a].ccIdx[k]; }
    }
  }
  else for( int k = 0; k < 3; k++ ) sl->setAlfEnabledFlag( ComponentID( k ), false );
  if( c->tools & SEAM_LMCS )
  {
    if( s == 0 ) { P->lmcsAps = std::make_shared<APS>(); P->lmcsAps->setAPSId( 0 ); P->lmcsAps->setAPSType( LMCS_APS ); }
    SliceReshapeInfo& si = P->lmcsAps->getReshaperAPSInfo();
    si.sliceReshaperEnableFlag = true; si.sliceReshaperModelPresentFlag = true; si.enableChromaAdj = c->lmcsChromaAdj;
    si.reshaperModelMinBinIdx = c->lmcsMinBin; si.reshaperModelMaxBinIdx = c->lmcsMaxBin; si.chrResScalingOffset = c->lmcsChrOffset;
    for( int i = 0; i < PIC_CODE_CW_BINS; i++ ) si.reshaperModelBinCWDelta[i] = c->lmcsDeltaCW[i];
    cur.ph->setLmcsEnabledFlag( true ); cur.ph->setLmcsChromaResidualScaleFlag( c->lmcsChromaAdj != 0 ); cur.ph->setLmcsAPS( P->lmcsAps );
    sl->setLmcsEnabledFlag( true );
  }
  }   // slices
  // the coding tree of every CTU
  SEAM_TR( "build: state ready, generating %u CTUs\n", pcv.sizeInCtus );
  Gen gen( cur, cur.pic.slices[0], *c );
  for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) { SEAM_TR( "ctu %u\n", a ); gen.sl = cur.pic.slices[sliceOfCtu( a )]; gen.ctu( a ); }
  // SAO merge syntax the way the parser leaves it (CABACReader.cpp:249-390: every component SAO_MODE_MERGE, typeIdc = SAO_MERGE_LEFT / _ABOVE), for about a
  // third of the CTUs whose neighbour is a merge candidate (same slice and tile: the first CU's left / above pointer, SampleAdaptiveOffset.cpp:573-620)
  if( doSao )
  {
    std::mt19937 mrng( c->seed * 7919u + 13u );
    for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
    {
      const CodingUnit* cu = cs.getCtuData( a ).cuPtr[CH_L][0];
      if( !cu || mrng() % 3 ) continue;
      const bool left = a % pcv.widthInCtus && cu->left, above = a >= pcv.widthInCtus && cu->above;
      if( !left && !above ) continue;
      const int type = left && ( !above || ( mrng() & 1 ) ) ? SAO_MERGE_LEFT : SAO_MERGE_ABOVE;
      SAOBlkParam& bp = cs.getCtuData( a ).saoParam; bp.reset();
      for( int k = 0; k < 3; k++ ) { bp[k].modeIdc = SAO_MODE_MERGE; bp[k].typeIdc = type; }
    }
  }
  cur.pic.progress = Picture::parsed;
  cur.pic.parseDone.unlock();
  return P.release();
}

struct StockCtx { std::unique_ptr<ThreadPool> pool; std::unique_ptr<DecLibRecon> rec; int threads = -1; };
static StockCtx& stockCtx( int threads )
{
  static StockCtx S;
  if( S.threads != threads )
  {
    if( S.rec ) S.rec->destroy();
    S.rec.reset(); S.pool.reset();
    S.pool.reset( new ThreadPool( threads, "seam" ) ); S.rec.reset( new DecLibRecon ); S.rec->create( S.pool.get(), 0, false ); S.threads = threads;
  }
  return S;
}
struct B200Ctx { std::unique_ptr<ThreadPool> pool; std::unique_ptr<b200glue::DecLibReconB200> rec; int threads = -1; bool dry = false; b200_geom g{}; };
static B200Ctx& b200Ctx( int threads, bool dry, const b200_geom& g )
{
  static B200Ctx S;
  if( S.threads != threads || S.dry != dry || memcmp( &S.g, &g, sizeof( g ) ) )      // a device context serves one picture geometry (a new sequence = a new decoder)
  {
    S.g = g;
    if( S.rec ) S.rec->destroy();
    S.rec.reset(); S.pool.reset();
    S.pool.reset( new ThreadPool( threads, "seamB200" ) ); S.rec.reset( new b200glue::DecLibReconB200 ); S.rec->create( S.pool.get(), 0, false ); S.rec->setDryRun( dry ); S.threads = threads; S.dry = dry;
  }
  return S;
}

static void readOut( Pic& P, int16_t* const out[3], uint8_t* colMotion, size_t colBytes )
{
  FakePicture& cur = *P.cur;
  cur.pic.cs->rebindPicBufs();
  if( out ) cur.getPlanes( P.g, out );
  if( colMotion )
  {
    const size_t n = std::min( colBytes, (size_t) cur.pic.cs->m_colMiMapSize * sizeof( ColocatedMotionInfo ) );
    memcpy( colMotion, cur.pic.cs->m_colMiMap.get(), n );
  }
}

}   // namespace seam

extern "C" void* ref_seam_create( const b200_geom* g, const ref_seam_cfg* cfg, const int16_t* const* refs, const b200_picture* filt )
{
  try { return seam::build( g, cfg, refs, filt ); }
  catch( std::exception& e ) { fprintf( stderr, "ref_seam_create: %s\n", e.what() ); return nullptr; }
}
extern "C" void* ref_seam_create_chained( const b200_geom* g, const ref_seam_cfg* cfg, const int16_t* const* refs, const b200_picture* filt, void* prev )
{
  try { return seam::build( g, cfg, refs, filt, static_cast<seam::Pic*>( prev ) ); }
  catch( std::exception& e ) { fprintf( stderr, "ref_seam_create_chained: %s\n", e.what() ); return nullptr; }
}
extern "C" void ref_seam_destroy( void* h ) { delete static_cast<seam::Pic*>( h ); }

extern "C" size_t ref_seam_col_motion_bytes( void* h )
{
  seam::Pic& P = *static_cast<seam::Pic*>( h );
  return (size_t) P.cur->pic.cs->m_colMiMapSize * sizeof( ColocatedMotionInfo );
}

// counts of what the generator produced: [0] CUs, [1] intra CUs, [2] TUs, [3] skip, [4] merge, [5] affine, [6] geo, [7] ciip, [8] mmvd, [9] CUs with residual, [10] sbt, [11] lfnst, [12] mts, [13] isp, [14] mip, [15] chroma-tree CUs
extern "C" void ref_seam_stats( void* h, int32_t st[16] )
{
  seam::Pic& P = *static_cast<seam::Pic*>( h );
  CodingStructure& cs = *P.cur->pic.cs;
  memset( st, 0, 16 * sizeof( int32_t ) );
  for( unsigned a = 0; a < cs.pcv->sizeInCtus; a++ )
    for( auto& cu : cs.traverseCUs( a ) )
    {
      st[0]++; st[1] += CU::isIntra( cu ); st[3] += cu.skip(); st[4] += cu.mergeFlag(); st[5] += cu.affineFlag(); st[6] += cu.geoFlag(); st[7] += cu.ciipFlag(); st[8] += cu.mmvdFlag();
      st[9] += cu.rootCbf(); st[10] += cu.sbtInfo() != 0; st[11] += cu.lfnstIdx() != 0; st[13] += cu.ispMode() != 0; st[14] += cu.mipFlag(); st[15] += isChroma( cu.chType() );
      for( auto& tu : TUTraverser( &cu.firstTU, cu.lastTU->next ) ) { st[2]++; st[12] += tu.mtsIdx( COMPONENT_Y ) > MTS_SKIP; }
    }
}

// (i) the reference's own DecLibRecon.  Returns the seconds between decompressPicture() and the return of waitForPrevDecompressedPic(), < 0 on error.
extern "C" double ref_seam_run_stock( void* h, int threads, int16_t* const out[3], uint8_t* colMotion, size_t colBytes )
{
  seam::Pic& P = *static_cast<seam::Pic*>( h );
  try
  {
    seam::StockCtx& S = seam::stockCtx( threads );
    Picture* pic = &P.cur->pic;
    if( P.ref[0] ) P.ref[0]->pic.borderExtStarted = false;      // every reference is extended once in its life: one of the four inside each run (the extension is idempotent)
    const auto t0 = std::chrono::steady_clock::now();
    S.rec->decompressPicture( pic );
    Picture* done = S.rec->waitForPrevDecompressedPic();
    const double secs = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
    if( done != pic || pic->error || pic->reconDone.hasException() )
    {
      try { pic->reconDone.checkAndRethrowException(); } catch( std::exception& e ) { fprintf( stderr, "ref_seam_run_stock: %s\n", e.what() ); }
      pic->reconDone.clearException();
      return -1.0;
    }
    seam::readOut( P, out, colMotion, colBytes );
    return secs;
  }
  catch( std::exception& e ) { fprintf( stderr, "ref_seam_run_stock: %s\n", e.what() ); return -2.0; }
}

// (ii) the GPU back end behind the same seam.  dry != 0: host stages only (no device needed), the flattened work lists come back through *flat
// (pointers into the recon object, valid until the next call) — this is how the CPU tests chain the flattener into the oracle.
extern "C" double ref_seam_run_b200( void* h, int threads, int dry, int16_t* const out[3], uint8_t* colMotion, size_t colBytes, b200_picture* flat )
{
  seam::Pic& P = *static_cast<seam::Pic*>( h );
  try
  {
    seam::B200Ctx& S = seam::b200Ctx( threads, dry != 0, P.g );
    { static const void* lastRefs = nullptr; if( lastRefs != (const void*) P.refs.get() ) { S.rec->resetDpb(); lastRefs = P.refs.get(); } }   // another reference set: its predecessor's Pictures are gone
    Picture* pic = &P.cur->pic;
    const auto t0 = std::chrono::steady_clock::now();
    try { S.rec->decompressPicture( pic ); }
    catch( ... ) { pic->reconDone.setException( std::current_exception() ); pic->error = true; }          // DecLib::reconPicture (DecLib.cpp:618-626)
    Picture* done = S.rec->waitForPrevDecompressedPic();
    const double secs = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
    S.rec->releasePicture( pic );                                        // the picture dies with its handle; the shared references stay resident on the device
    if( done != pic || pic->error || pic->reconDone.hasException() )
    {
      int rc = -1;
      try { pic->reconDone.checkAndRethrowException(); }
      catch( UnsupportedFeatureException& e ) { fprintf( stderr, "ref_seam_run_b200: unsupported: %s\n", e.what() ); rc = -4; }
      catch( std::exception& e ) { fprintf( stderr, "ref_seam_run_b200: %s\n", e.what() ); }
      pic->reconDone.clearException();
      return rc;
    }
    if( getenv( "SEAM_TIMING" ) ) { double cm[3]; S.rec->cpuStageMs( cm ); fprintf( stderr, "b200 host CPU ms summed over threads: MIDER %.2f | Bs + grid %.2f | CU / TU walk %.2f\n", cm[0], cm[1], cm[2] ); }
    if( getenv( "SEAM_TIMING" ) ) { const double* st = S.rec->stageTimes(); fprintf( stderr, "b200 host stages [ms]: tables %.2f | mider %.2f | flatten %.2f | submit %.2f | joined %.2f | finish %.2f\n", st[0] * 1e3, ( st[1] - st[0] ) * 1e3, ( st[2] - st[1] ) * 1e3, ( st[3] - st[2] ) * 1e3, ( st[4] - st[3] ) * 1e3, ( st[5] - st[4] ) * 1e3 ); }
    if( flat ) *flat = S.rec->flattened();
    if( !dry ) seam::readOut( P, out, colMotion, colBytes );
    else if( colMotion ) seam::readOut( P, nullptr, colMotion, colBytes );      // dry run: the motion field with zero DMVR deltas
    return secs;
  }
  catch( std::exception& e ) { fprintf( stderr, "ref_seam_run_b200: %s\n", e.what() ); return -2.0; }
}

namespace seam { static std::vector<std::unique_ptr<b200glue::DecLibReconB200>>& pipeDev() { static std::vector<std::unique_ptr<b200glue::DecLibReconB200>> v; return v; } }
// the work lists instance k of the last ref_seam_run_pipelined( backend 1 / 2 ) call flattened last (pointers into the recon object)
extern "C" int ref_seam_pipelined_flat( int k, b200_picture* flat )
{
  auto& dev = seam::pipeDev();
  if( k < 0 || k >= (int) dev.size() || !flat ) return -1;
  *flat = dev[k]->flattened(); return 0;
}

// (iii) Either back end the way DecLib drives it (DecLib.h:70, DecLib.cpp:560-640): `depth` recon instances on one thread pool take the pictures in turn;
// an instance is handed its next picture as soon as its previous one has been waited for, so up to `depth` pictures are in flight.  Returns the seconds
// for all n pictures (no output read-back), < 0 on error.  backend 0: the reference's DecLibRecon, 1: DecLibReconB200, 2: DecLibReconB200 in dry-run mode (host stages only).
// ref_seam_read_out() fetches a picture's planes / motion field afterwards.
// the DecLibReconB200 instances of ref_seam_run_pipelined() complete their pictures in a pool task of their own (setAsyncFinish) from the next run on
extern "C" void ref_seam_set_async_finish( int on ) { b200glue::DecLibReconB200::asyncFinishDefault() = on != 0; }

extern "C" double ref_seam_run_pipelined( void* const* hs, int n, int threads, int backend, int depth )
{
  if( n <= 0 || depth < 1 || depth > 4 ) return -3.0;
  try
  {
    seam::Pic& P0 = *static_cast<seam::Pic*>( hs[0] );
    static std::unique_ptr<ThreadPool> pool; static int poolThreads = -1;
    static std::vector<std::unique_ptr<DecLibRecon>> stock; auto& dev = seam::pipeDev(); static b200_geom devGeom{};
    if( poolThreads != threads || ( backend >= 1 && memcmp( &devGeom, &P0.g, sizeof( devGeom ) ) ) )
    {
      for( auto& r : stock ) r->destroy(); for( auto& r : dev ) r->destroy();
      stock.clear(); dev.clear(); pool.reset( new ThreadPool( threads, "seamPipe" ) ); poolThreads = threads; devGeom = P0.g;
    }
    while( backend == 0 && (int) stock.size() < depth ) { stock.emplace_back( new DecLibRecon ); stock.back()->create( pool.get(), (unsigned) stock.size() - 1, false ); }
    while( backend >= 1 && (int) dev.size() < depth )   { dev.emplace_back( new b200glue::DecLibReconB200 ); dev.back()->create( pool.get(), (unsigned) dev.size() - 1, false ); }
    if( backend >= 1 ) { for( auto& r : dev ) { r->setDryRun( backend == 2 ); r->setAsyncFinish( b200glue::DecLibReconB200::asyncFinishDefault() ); } dev[0]->resetDpb(); }
    bool bad = false; std::vector<Picture*> pendingRelease;
    auto waitOne = [&]( int k )
    {
      Picture* done = backend ? dev[k]->waitForPrevDecompressedPic() : stock[k]->waitForPrevDecompressedPic();
      if( !done ) return;
      if( done->error || done->reconDone.hasException() ) { bad = true; done->reconDone.clearException(); }
      // the slot of a finished picture is given back once no picture in flight can still wait on it: a picture handed over earlier may reference it and not have been
      // submitted yet (its submission polls the slot's state, and a NEW picture taking the slot over would look like a reference in flight).  DecLib's picture list
      // keeps such a picture referenced (stillReferenced); here: released `depth` pictures later.
      if( backend ) { pendingRelease.push_back( done ); while( (int) pendingRelease.size() > depth ) { dev[0]->releasePicture( pendingRelease.front() ); pendingRelease.erase( pendingRelease.begin() ); } }
    };
    const auto t0 = std::chrono::steady_clock::now();
    for( int i = 0; i < n; i++ )
    {
      const int k = i % depth;
      waitOne( k );
      Picture* pic = &static_cast<seam::Pic*>( hs[i] )->cur->pic;
      try { if( backend ) dev[k]->decompressPicture( pic ); else stock[k]->decompressPicture( pic ); }
      catch( ... ) { pic->reconDone.setException( std::current_exception() ); pic->error = true; }
    }
    for( int j = 0; j < depth; j++ ) waitOne( ( n + j ) % depth );
    if( backend ) for( Picture* p : pendingRelease ) dev[0]->releasePicture( p );
    const double secs = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
    return bad ? -1.0 : secs;
  }
  catch( std::exception& e ) { fprintf( stderr, "ref_seam_run_pipelined: %s\n", e.what() ); return -2.0; }
}
extern "C" void ref_seam_read_out( void* h, int16_t* const out[3], uint8_t* colMotion, size_t colBytes ) { seam::readOut( *static_cast<seam::Pic*>( h ), out, colMotion, colBytes ); }
