// DecLibReconB200.h — the drop-in for DecLibRecon (reference DecoderLib/DecLibRecon.h:143-200): same five methods, reconstruction on a B200
// through the C ABI of include/vvdec_b200.h.  Header-only glue that lives INSIDE a VVdeC build; it is compiled against
// the reference headers by oracle/Makefile.ref (it is included by oracle/ref_shim.cpp), its building blocks — flattenTU, flattenPU /
// flattenSbTmvp, flattenSAO / flattenALF / buildAlfTables / flattenLfCtu — are pinned one by one against the reference
// (tests/test_k1_oracle_vs_ref.py, test_flatten_pu_vs_ref.py, test_flatten_filters_vs_ref.py).  What cannot be exercised on the build box is
// the walk over a *parsed* picture: there is no VVC bitstream or encoder here.
//
// The CPU keeps: parsing (DecLibParser / DecSlice), motion derivation (DecCu::TaskDeriveCtuMotionInfo, DecCu.cpp:62), boundary strengths
// (LoopFilter::calcFilterStrengthsCTU, LoopFilter.cpp:360), TaskFinishMotionInfo (DecCu.cpp:161).  Pictures that use a tool the device path
// does not have (ISP intra blocks, IBC CUs, RPR, wrap-around, sub-picture clipping, virtual-boundary ALF) throw UnsupportedFeatureException; a
// deployment keeps a stock DecLibRecon next to this class and routes those pictures to it.
#pragma once
#include <vector>
#include <map>
#include "vvdec_b200.h"
#include "flatten_tu.h"
#include "flatten_pu.h"
#include "flatten_filters.h"
#include "flatten_intra.h"
#include "CommonLib/Picture.h"
#include "CommonLib/LoopFilter.h"
#include "CommonLib/Reshape.h"
#include "CommonLib/WeightPrediction.h"
#include "DecoderLib/DecCu.h"

namespace b200glue
{
using namespace vvdec;

// grow-only pinned host array (cudaHostAlloc through the C ABI's registration call: memory stays owned by the vector)
template<class T> struct PinnedVec
{
  std::vector<T> v; const void* reg = nullptr; size_t regBytes = 0;
  void clear() { v.clear(); }
  void pin() { if( v.capacity() * sizeof( T ) != regBytes || (const void*) v.data() != reg ) { if( reg ) b200_host_unregister( const_cast<void*>( reg ) ); reg = v.data(); regBytes = v.capacity() * sizeof( T ); if( regBytes ) b200_host_register( const_cast<void*>( reg ), regBytes ); } }
  ~PinnedVec() { if( reg ) b200_host_unregister( const_cast<void*>( reg ) ); }
};

class DecLibReconB200
{
  b200_ctx*  m_ctx = nullptr;
  b200_geom  m_geom{};
  int        m_numSlots = 0, m_arena = -1;
  Picture*   m_currDecompPic = nullptr;
  std::map<const Picture*, int> m_slotOf;          // DPB slot of every picture the device holds
  std::vector<const Picture*>   m_slotOwner;
  // per-picture work lists (pinned)
  PinnedVec<b200_pu> m_pus; PinnedVec<b200_tu> m_tus; PinnedVec<int16_t> m_coefs; PinnedVec<b200_intra_tu> m_intra;
  PinnedVec<b200_lf_param> m_lf[2]; PinnedVec<b200_sao_ctu> m_sao; PinnedVec<b200_alf_ctu> m_alf; PinnedVec<b200_lmcs_vpdu> m_vpdus;
  PinnedVec<int32_t> m_dmvr;
  std::vector<b200_wp> m_wp; std::map<std::pair<int, int>, int> m_wpIdx;
  AlfTableStore m_alfStore; b200_alf_tables m_alfTabs{}; b200_lmcs m_lmcs{}; b200_vb m_vb{}; b200_lf_slice m_lfSlice{}; b200_lf_seq m_lfSeq{};
  // the CPU stages that stay
  std::vector<MotionInfo> m_motionInfo; std::vector<LoopFilterParam> m_loopFilterParam; std::vector<Mv> m_dmvrMvCache;
  LoopFilter m_cLoopFilter; SampleAdaptiveOffset m_cSAO; AdaptiveLoopFilter m_cALF; Reshape m_cReshaper; DecCu m_cCuDecoder; TrQuant* m_trQuant = nullptr;
  PelStorage m_fltBuf;

  static void check( int rc ) { if( rc == B200_ERR_UNSUPPORTED ) THROW_UNSUPPORTED( b200_last_error() ); if( rc == B200_ERR_PARAM ) THROW_RECOVERABLE( b200_last_error() ); if( rc < 0 ) THROW_FATAL( b200_last_error() ); }

  int slotFor( const Picture* pic )
  {
    auto it = m_slotOf.find( pic ); if( it != m_slotOf.end() ) return it->second;
    for( int s = 0; s < m_numSlots; s++ ) if( !m_slotOwner[s] || !m_slotOwner[s]->stillReferenced ) { if( m_slotOwner[s] ) m_slotOf.erase( m_slotOwner[s] ); m_slotOwner[s] = pic; m_slotOf[pic] = s; return s; }
    THROW_FATAL( "DecLibReconB200: device DPB is full" );
  }

public:
  void create( TrQuant* trQuant, int dpbSlots ) { m_trQuant = trQuant; m_numSlots = dpbSlots; m_slotOwner.assign( dpbSlots, nullptr ); }
  void destroy() { if( m_ctx ) b200_ctx_destroy( m_ctx ); m_ctx = nullptr; }
  Picture* getCurrPic() const { return m_currDecompPic; }

  // DecLibRecon::decompressPicture (DecLibRecon.cpp:429): flatten the parsed picture and hand it to the device; returns without waiting.
  void decompressPicture( Picture* pic )
  {
    m_currDecompPic = pic;
    CodingStructure& cs = *pic->cs;
    const SPS& sps = *cs.sps; const PPS& pps = *cs.pps; const PreCalcValues& pcv = *cs.pcv;
    pic->progress = Picture::reconstructing;
    if( !m_ctx )
    {
      m_geom.width = pcv.lumaWidth; m_geom.height = pcv.lumaHeight; m_geom.bitDepth = sps.getBitDepth(); m_geom.chromaFormat = sps.getChromaFormatIdc() == CHROMA_420 ? 1 : 0;
      m_geom.ctuSize = pcv.maxCUWidth; m_geom.stride[0] = pcv.lumaWidth; m_geom.stride[1] = m_geom.stride[2] = pcv.lumaWidth >> 1;
      if( sps.getChromaFormatIdc() != CHROMA_420 && sps.getChromaFormatIdc() != CHROMA_400 ) THROW_UNSUPPORTED( "DecLibReconB200: 4:2:0 and 4:0:0 only" );
      check( b200_ctx_create( &m_ctx, &m_geom, m_numSlots, 4, -1 ) );
    }
    if( sps.getUseWrapAround() || pic->subPictures.size() > 1 || cs.picHeader->getVirtualBoundariesPresentFlag() ) THROW_UNSUPPORTED( "DecLibReconB200: wrap-around / sub-pictures / virtual boundaries" );
    pic->parseDone.wait();

    // ---- CPU stages per CTU, in the order of ctuTask's MIDER / LF_INIT cases (DecLibRecon.cpp:762-829) ----
    m_motionInfo.resize( (size_t) pcv.num4x4CtuBlks * pcv.sizeInCtus ); m_loopFilterParam.assign( (size_t) pcv.num4x4CtuBlks * pcv.sizeInCtus * 2, LoopFilterParam{} );
    m_dmvrMvCache.assign( (size_t) pcv.num8x8CtuBlks * pcv.sizeInCtus, Mv() ); cs.m_dmvrMvCache = m_dmvrMvCache.data();
    std::vector<MotionHist> hist( pcv.heightInCtus );
    for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
    {
      CtuData& cd = cs.getCtuData( a );
      cd.motion = &m_motionInfo[(size_t) pcv.num4x4CtuBlks * a];
      const UnitArea ctuArea = getCtuArea( cs, a % pcv.widthInCtus, a / pcv.widthInCtus, true );
      if( !cd.slice->isIntra() ) m_cCuDecoder.TaskDeriveCtuMotionInfo( cs, a, ctuArea, hist[a / pcv.widthInCtus] );            // ctuTask MIDER (DecLibRecon.cpp:762-781)
    }
    for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
    {
      CtuData& cd = cs.getCtuData( a );
      cd.lfParam[0] = &m_loopFilterParam[(size_t) pcv.num4x4CtuBlks * ( 2 * a )]; cd.lfParam[1] = &m_loopFilterParam[(size_t) pcv.num4x4CtuBlks * ( 2 * a + 1 )];
      m_cLoopFilter.calcFilterStrengthsCTU( cs, a );
    }

    // ---- reference pictures -> device DPB slots, explicit weights ----
    const Slice& slice0 = *pic->slices[0];
    SlotMap sm; memset( &sm, -1, sizeof( sm ) );
    for( int l = 0; l < 2; l++ ) for( int i = 0; i < slice0.getNumRefIdx( RefPicList( l ) ); i++ ) sm.slot[l][i] = (int8_t) slotFor( slice0.getRefPic( RefPicList( l ), i ) );
    if( pic->slices.size() > 1 ) THROW_UNSUPPORTED( "DecLibReconB200: one slice per picture (reference lists and ALF / LMCS tables are per slice)" );
    m_wp.clear(); m_wpIdx.clear();
    auto wpIdxOf = [&]( int r0, int r1 ) -> int
    {
      auto it = m_wpIdx.find( { r0, r1 } ); if( it != m_wpIdx.end() ) return it->second;
      WPScalingParam w0[3], w1[3]; WeightPrediction wpObj; wpObj.getWpScaling( &slice0, r0, r1, w0, w1 );
      b200_wp e{}; const bool bi = r0 >= 0 && r1 >= 0; const WPScalingParam* u = r0 >= 0 ? w0 : w1;
      for( int c = 0; c < 3; c++ ) { e.w0[c] = bi ? w0[c].w : u[c].w; e.w1[c] = bi ? w1[c].w : 0; e.offset[c] = bi ? w0[c].offset : u[c].offset; e.shift[c] = bi ? w0[c].shift : u[c].shift; }
      m_wp.push_back( e ); return m_wpIdx[{ r0, r1 }] = (int) m_wp.size();
    };

    // ---- flatten CUs / TUs (TaskTrafoCtu + TaskInterCtu walks, DecCu.cpp:106-134) ----
    m_pus.clear(); m_tus.clear(); m_coefs.clear(); m_intra.clear();
    for( unsigned a = 0; a < pcv.sizeInCtus; a++ )
      for( auto& cu : cs.traverseCUs( a ) )
      {
        if( CU::isIntra( cu ) )
        {
          // K6: regular intra modes on the device.  One b200_intra_tu per TU component in decoding order (the order DecCu::predAndReco walks them,
          // DecCu.cpp:284-288); the residual of a coded component goes through K1 into the residual planes (B200_TU_RESI) and is added by K6.
          if( sps.getUseReshaper() && cs.picHeader->getLmcsEnabledFlag() ) THROW_UNSUPPORTED( "DecLibReconB200: intra CUs with LMCS" );
          for( auto& tu : TUTraverser( &cu.firstTU, cu.lastTU->next ) )
            for( const CompArea& area : tu.blocks )
            {
              if( !area.valid() ) continue;
              b200_intra_tu ir;
              if( flattenIntraTU( tu, area.compID(), ir ) != FLATTEN_INTRA_OK ) THROW_UNSUPPORTED( "DecLibReconB200: ISP / ACT intra block (SURVEY 8f-1)" );
              b200_tu r;
              if( flattenTU( tu, area.compID(), *m_trQuant, m_coefs.v, r ) ) { r.flags |= B200_TU_RESI; m_tus.v.push_back( r ); }
              if( TU::getCbf( tu, area.compID() ) || ( isChroma( area.compID() ) && tu.jointCbCr ) ) ir.flags |= B200_INTRA_ADD_RESI;      // DecCu.cpp:390
              m_intra.v.push_back( ir );
            }
          continue;
        }
        if( !CU::isInter( cu ) ) THROW_UNSUPPORTED( "DecLibReconB200: IBC CU (SURVEY 8f-1)" );
        bool ciipComp[3] = { false, false, false };
        if( cu.ciipFlag() )
        {
          // CIIP: the inter prediction comes from K2 like any merge CU; K6 blends a planar intra block into it (predBlendIntraCiip) and adds the residual
          if( sps.getUseReshaper() && cs.picHeader->getLmcsEnabledFlag() ) THROW_UNSUPPORTED( "DecLibReconB200: CIIP CUs with LMCS" );
          for( int c = 0; c < (int) getNumberValidComponents( cu.chromaFormat ); c++ )
          {
            b200_intra_tu ir;
            if( flattenCiipBlock( cu, ComponentID( c ), ir ) != FLATTEN_INTRA_OK ) continue;
            if( TU::getCbf( cu.firstTU, ComponentID( c ) ) || ( c && cu.firstTU.jointCbCr ) ) ir.flags |= B200_INTRA_ADD_RESI;
            m_intra.v.push_back( ir ); ciipComp[c] = true;
          }
        }
        FlattenPuResult rc;
        if( cu.mergeType() == MRG_TYPE_SUBPU_ATMVP ) rc = flattenSbTmvp( cu, sm, wpIdxOf, [&]( const b200_pu& r ) { m_pus.v.push_back( r ); } );
        else { b200_pu r; rc = flattenPU( cu, sm, wpIdxOf, r ); if( rc == FLATTEN_PU_OK ) m_pus.v.push_back( r ); }
        if( rc != FLATTEN_PU_OK ) THROW_UNSUPPORTED( "DecLibReconB200: inter tool outside the device path" );
        if( cu.rootCbf() )
          for( auto& tu : TUTraverser( &cu.firstTU, cu.lastTU->next ) )
            for( int c = 0; c < (int) getNumberValidComponents( cu.chromaFormat ); c++ ) { b200_tu r; if( flattenTU( tu, ComponentID( c ), *m_trQuant, m_coefs.v, r ) ) { if( ciipComp[c] ) r.flags |= B200_TU_RESI; m_tus.v.push_back( r ); } }
      }

    // ---- in-loop filter parameters ----
    const int W4 = ( pcv.lumaWidth + 3 ) >> 2, H4 = ( pcv.lumaHeight + 3 ) >> 2;
    for( int d = 0; d < 2; d++ ) { m_lf[d].v.assign( (size_t) W4 * H4, b200_lf_param{} ); for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) flattenLfCtu( cs, a, d, m_lf[d].v.data() ); }
    m_lfSlice = b200_lf_slice{}; m_lfSlice.disable = slice0.getDeblockingFilterDisable();
    m_lfSlice.betaOffsetDiv2[0] = slice0.getDeblockingFilterBetaOffsetDiv2(); m_lfSlice.tcOffsetDiv2[0] = slice0.getDeblockingFilterTcOffsetDiv2();
    m_lfSlice.betaOffsetDiv2[1] = slice0.getDeblockingFilterCbBetaOffsetDiv2(); m_lfSlice.tcOffsetDiv2[1] = slice0.getDeblockingFilterCbTcOffsetDiv2();
    m_lfSlice.betaOffsetDiv2[2] = slice0.getDeblockingFilterCrBetaOffsetDiv2(); m_lfSlice.tcOffsetDiv2[2] = slice0.getDeblockingFilterCrTcOffsetDiv2();
    const bool doSao = sps.getUseSAO(), doAlf = sps.getUseALF() && ( slice0.getAlfEnabledFlag( COMPONENT_Y ) || slice0.getAlfEnabledFlag( COMPONENT_Cb ) || slice0.getAlfEnabledFlag( COMPONENT_Cr ) );
    m_sao.v.assign( pcv.sizeInCtus, b200_sao_ctu{} ); m_alf.v.assign( pcv.sizeInCtus, b200_alf_ctu{} );
    for( unsigned a = 0; a < pcv.sizeInCtus && doSao; a++ )
    {
      bool av[8];
      m_cSAO.deriveLoopFilterBoundaryAvailibility( cs, Position( ( a % pcv.widthInCtus ) * pcv.maxCUWidth, ( a / pcv.widthInCtus ) * pcv.maxCUHeight ), av[0], av[1], av[2], av[3], av[4], av[5], av[6], av[7] );
      flattenSAO( cs.getCtuData( a ).saoParam, av, getNumberValidComponents( pcv.chrFormat ), m_sao.v[a] );     // saoParam after reconstructBlkSAOParam (parser side)
    }
    if( doAlf )
    {
      for( unsigned a = 0; a < pcv.sizeInCtus; a++ ) flattenALF( cs.getCtuData( a ).alfParam, m_alf.v[a] );
      m_alfTabs = buildAlfTables( slice0, &m_cALF.m_fixedFilterSetCoeffDec[0][0], m_cALF.m_clipDefault, m_alfStore );
    }
    // ---- LMCS ----
    const bool doLmcs = sps.getUseReshaper() && cs.picHeader->getLmcsEnabledFlag() && slice0.getLmcsEnabledFlag();
    if( doLmcs )
    {
      m_cReshaper.createDec( sps.getBitDepth() ); m_cReshaper.initSlice( slice0.getNalUnitLayerId(), *slice0.getPicHeader(), slice0.getVPS_nothrow() );
      m_lmcs = b200_lmcs{}; m_lmcs.chromaAdj = m_cReshaper.m_sliceReshapeInfo.enableChromaAdj;
      m_lmcs.minBinIdx = m_cReshaper.m_sliceReshapeInfo.reshaperModelMinBinIdx; m_lmcs.maxBinIdx = m_cReshaper.m_sliceReshapeInfo.reshaperModelMaxBinIdx; m_lmcs.orgCW = m_cReshaper.m_initCW;
      for( int i = 0; i < 17; i++ ) { m_lmcs.reshapePivot[i] = m_cReshaper.m_reshapePivot[i]; m_lmcs.inputPivot[i] = m_cReshaper.m_inputPivot[i]; }
      for( int i = 0; i < 16; i++ ) { m_lmcs.fwdScaleCoef[i] = m_cReshaper.m_fwdScaleCoef[i]; m_lmcs.chromaAdjHelpLUT[i] = m_cReshaper.m_chromaAdjHelpLUT[i]; }
      m_lmcs.invLUT = m_cReshaper.m_invLUT;
      const int vs = pcv.maxCUWidth == 128 ? 64 : pcv.maxCUWidth, vW = ( pcv.lumaWidth + vs - 1 ) / vs, vH = ( pcv.lumaHeight + vs - 1 ) / vs;
      m_vpdus.v.assign( (size_t) vW * vH, b200_lmcs_vpdu{} );
      for( int j = 0; j < vH; j++ ) for( int i = 0; i < vW; i++ )
      {
        const Position tl( i * vs, j * vs );                                                     // Reshape.cpp:217-219
        const CodingUnit* cu = cs.getCU( tl, CHANNEL_TYPE_LUMA );
        const CodingUnit* above = cs.getCURestricted( cu->lumaPos().offset( 0, -1 ), *cu, CHANNEL_TYPE_LUMA, cu->ly() == tl.y ? cu : cu->above );
        const CodingUnit* left  = cs.getCURestricted( cu->lumaPos().offset( -1, 0 ), *cu, CHANNEL_TYPE_LUMA, cu->lx() == tl.x ? cu : cu->left );
        m_vpdus.v[(size_t) j * vW + i] = b200_lmcs_vpdu{ (uint16_t) cu->lx(), (uint16_t) cu->ly(), (uint8_t) ( left != nullptr ), (uint8_t) ( above != nullptr ) };
      }
      m_lmcs.vpdus = m_vpdus.v.data();
    }

    // ---- submit ----
    for( PinnedVec<b200_lf_param>& l : m_lf ) l.pin();
    m_pus.pin(); m_tus.pin(); m_coefs.pin(); m_sao.pin(); m_alf.pin(); m_intra.pin();
    b200_picture p{};
    p.dstSlot = slotFor( pic );
    p.flags = ( slice0.getDeblockingFilterDisable() ? 0 : B200_PIC_DEBLOCK ) | ( doSao ? B200_PIC_SAO : 0 ) | ( doAlf ? B200_PIC_ALF : 0 ) | ( doLmcs ? B200_PIC_LMCS : 0 );
    p.pus = m_pus.v.data(); p.numPus = m_pus.v.size(); p.numDmvr = m_dmvrMvCache.size();
    p.tus = m_tus.v.data(); p.numTus = m_tus.v.size(); p.coefs = m_coefs.v.data(); p.numCoefs = m_coefs.v.size();
    p.lfV = m_lf[0].v.data(); p.lfH = m_lf[1].v.data(); p.lfSlices = &m_lfSlice; p.numLfSlices = 1; p.lfSeq = &m_lfSeq;
    p.sao = m_sao.v.data(); p.vb = &m_vb; p.alf = m_alf.v.data(); p.alfTabs = &m_alfTabs;
    p.wp = m_wp.data(); p.numWp = (int32_t) m_wp.size(); p.lmcs = doLmcs ? &m_lmcs : nullptr;
    p.intraTus = m_intra.v.data(); p.numIntraTus = m_intra.v.size();
    m_arena = b200_decompress_picture( m_ctx, &p );
    check( m_arena );
  }

  // DecLibRecon::waitForPrevDecompressedPic (DecLibRecon.cpp:684): device done -> DMVR deltas -> TaskFinishMotionInfo -> output planes.
  Picture* waitForPrevDecompressedPic()
  {
    Picture* pic = m_currDecompPic; if( !pic ) return nullptr;
    CodingStructure& cs = *pic->cs; const PreCalcValues& pcv = *cs.pcv;
    m_dmvr.v.assign( m_dmvrMvCache.size() * 2, 0 ); m_dmvr.pin();
    check( b200_wait_picture( m_ctx, m_arena, m_dmvr.v.data(), m_dmvrMvCache.size() ) );
    for( size_t i = 0; i < m_dmvrMvCache.size(); i++ ) m_dmvrMvCache[i] = Mv( m_dmvr.v[2 * i], m_dmvr.v[2 * i + 1] );
    for( unsigned a = 0; a < pcv.sizeInCtus; a++ )                                                                                         // colMotion for later TMVP (DecCu.cpp:161),
      if( !cs.getCtuData( a ).slice->isIntra() && pic->stillReferenced ) m_cCuDecoder.TaskFinishMotionInfo( cs, a, a % pcv.widthInCtus, a / pcv.widthInCtus );   // under ctuTask's conditions (DecLibRecon.cpp:860-867)
    if( pic->neededForOutput )
    {
      int16_t* planes[3] = { cs.getRecoBuf( COMPONENT_Y ).buf, nullptr, nullptr };
      if( m_geom.chromaFormat ) { planes[1] = cs.getRecoBuf( COMPONENT_Cb ).buf; planes[2] = cs.getRecoBuf( COMPONENT_Cr ).buf; }
      if( cs.getRecoBuf( COMPONENT_Y ).stride != (ptrdiff_t) m_geom.stride[0] ) THROW_UNSUPPORTED( "DecLibReconB200: output planes must be allocated without margins (vvdec_decoder_open_with_allocator)" );
      check( b200_get_frame( m_ctx, m_slotOf[pic], planes ) );
    }
    pic->progress = Picture::reconstructed;
    pic->reconDone.unlock();
    m_currDecompPic = nullptr;
    return pic;
  }
};

}   // namespace b200glue
