/* ref_shim.h — C view of oracle/_ref/libvvdec_ref.so (the UNMODIFIED reference + our extern "C" shim).
 * TEST INFRASTRUCTURE ONLY. `simd` = 0 selects the reference's scalar *Core functions, 1 the SIMD
 * re-bindings (what read_x86_extension_flags() picks on the host: SSE4.1 or AVX2). */
#ifndef REF_SHIM_H
#define REF_SHIM_H
#include <stdint.h>
#include <stddef.h>
#include "../include/vvdec_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

const char* ref_simd_level(void);

/* ---- K1, pointer level ---- */
void ref_dequant(int simd, int width, int maxX, int maxY, int scale, const int16_t* q, size_t qStride, int32_t* coef,
                 int rightShift, int inputMaximum, int32_t transformMaximum);
void ref_inv_lfnst(int32_t* src, int32_t* dst, unsigned set, unsigned index, unsigned size, int zeroOutSize);
void ref_inv_1d(int simd, int trType, int n, const int32_t* src, int32_t* dst, int shift, int line, int skipLine,
                int skipLine2, int clip, int32_t outMin, int32_t outMax);
void ref_cpy_resi_clip(int simd, const int32_t* src, int16_t* dst, ptrdiff_t stride, unsigned w, unsigned h,
                       int32_t outMin, int32_t outMax, int32_t round, int32_t shift);

/* ---- K1, TU level: builds a real vvdec TransformUnit/CodingUnit/SPS/PPS/Slice from syntax values, runs
 *      (a) our flattener (vvdec_b200/vvdec_glue/flatten_tu.h) and (b) the reference's own
 *      QpParam + TrQuant::invTransformNxN (+ invTransformICT) exactly as DecCu::reconstructResi does. ---- */
typedef struct ref_tu_syntax {
  int32_t w, h;            /* luma size of the CU == TU (no TU split)                 */
  int32_t comp;            /* component under test 0/1/2                              */
  int32_t bitDepth;
  int32_t predMode;        /* 0 inter, 1 intra (vvdec PredMode)                       */
  int32_t qp, chromaQpAdj; /* cu.qp, cu.chromaQpAdj                                   */
  int32_t cbQpOffset, crQpOffset, jointQpOffset;   /* pps offsets                     */
  int32_t depQuant;
  int32_t mtsIdx;          /* tu.mtsIdx(comp) 0..5                                    */
  int32_t lfnstIdx;        /* 0..2                                                    */
  int32_t intraDirL, intraDirC;
  int32_t mipFlag, ispMode, sbtIdx, sbtPos;
  int32_t bdpcmL, bdpcmC;
  int32_t jointCbCr;       /* 0..3 ; cbf bits derived                                  */
  int32_t jointCbCrSign;
  int32_t maxScanPosX, maxScanPosY;
  int32_t spsMTS, spsIntraMTS, spsInterMTS, spsLFNST;   /* implicit MTS = spsMTS && !spsIntraMTS (Slice.h:1796) */
  int32_t sepTree;         /* dual tree chroma CU                                      */
} ref_tu_syntax;

/* levels: compW*compH int16 row-major for the coded component. Outputs: resi0 = plane of the coded
 * component's area after the reference ran (compW*compH), resi1 = other chroma plane (joint CbCr) or
 * untouched; rec = our flattened record; coefs/numCoefs = packed corner. Returns 1 if a record was made. */
int ref_tu_case(const ref_tu_syntax* s, const int16_t* levels, int16_t* resi0, int16_t* resi1,
                b200_tu* rec, int16_t* coefs, int32_t* numCoefs);

/* ---- K3 deblocking ---- */
void ref_lf_pel_filter_luma(int simd, int16_t* src, ptrdiff_t step, ptrdiff_t offset, int tc, int sw, int thrCut,
                            int filterSecondP, int filterSecondQ, int bitDepth);
void ref_lf_filtering_pq(int simd, int16_t* src, ptrdiff_t step, ptrdiff_t offset, int numP, int numQ, int tc);
/* Picture level: builds a real vvdec Picture/CodingStructure (SPS/PPS/PicHeader/Slice via the reference's own
 * create/finalInit), copies planes + LoopFilterParam grids in, runs LoopFilter::loopFilterCTU(EDGE_VER) over all
 * CTUs then (EDGE_HOR) over all CTUs — the order the CTU state machine guarantees (DecLibRecon.cpp:943-989). */
int ref_lf_deblock_picture(int simd, const b200_geom* g, int16_t* const planes[3], const b200_lf_param* lfV,
                           const b200_lf_param* lfH, const uint8_t* ctuSlice, const b200_lf_slice* slices, int numSlices,
                           const b200_lf_seq* seq, int dirs);

/* ---- K4 SAO ---- */
void ref_sao_offset_block(int simd, int bitDepth, int typeIdx, const int* offset32, int startIdx /*BO first band*/, const int16_t* src, int16_t* dst,
                          ptrdiff_t srcStride, ptrdiff_t dstStride, int width, int height, unsigned avail,
                          int numVerVb, const int* verVb, int numHorVb, const int* horVb);
/* real SampleAdaptiveOffset::SAOProcessCTU over a real CodingStructure (one CU per CTU, single slice/tile) */
int ref_sao_picture(int simd, const b200_geom* g, const int16_t* const src[3], int16_t* const dst[3], const b200_sao_ctu* ctus, const b200_vb* vb);

/* ---- K5 ALF ---- */
void ref_alf_classify(int simd, uint16_t* cls /*64*/, const int16_t* srcLuma, ptrdiff_t stride, int planeW, int planeH,
                      int blkX, int blkY, int blkW, int blkH, int shift, int vbCtuHeight, int vbPos);
void ref_alf_filter_blk(int simd, int is7x7, const uint16_t* cls, int16_t* dst, ptrdiff_t dstStride, const int16_t* src, ptrdiff_t srcStride,
                        int planeW, int planeH, int blkX, int blkY, int blkW, int blkH, const int16_t* coeff, const int16_t* clip,
                        int bitDepth, int vbCtuHeight, int vbPos);
void ref_alf_ccalf_blk(int simd, int16_t* dstChroma, ptrdiff_t chromaStride, const int16_t* srcLuma, ptrdiff_t lumaStride,
                       int lumaW, int lumaH, int cX, int cY, int cW, int cH, const int16_t* coeff, int bitDepth, int vbCtuHeight, int vbPos);
/* real AdaptiveLoopFilter::prepareCTU + processCTU over a real CodingStructure; APS objects are built from tabs */
int ref_alf_picture(int simd, const b200_geom* g, const int16_t* const src[3], int16_t* const dst[3], const b200_alf_ctu* ctus,
                    const b200_alf_tables* tabs);

/* ---- K2 inter prediction: the real InterPrediction::motionCompensation on real CodingUnits ----
 * refs[slot*3+comp]: 4 reference pictures; list 0 = {slot 0 (POC 4), slot 1 (POC 0)}, list 1 = {slot 2 (POC 12), slot 3 (POC 16)},
 * current POC 8, so (slot 0, slot 2) and (slot 1, slot 3) are the equal-distance pairs BDOF / DMVR require.
 * Each PU becomes a CU (merge flags etc. chosen so that the reference itself derives the same bio/dmvr decision as pu->flags;
 * returns -1 if it did not). pred = the CU's prediction (written to dst planes), dmvrMv = m_dmvrMvCache. */
int ref_mc_predict(int simd, const b200_geom* g, int16_t* const dst[3], const int16_t* const* refs, const b200_pu* pus, size_t numPus,
                   int32_t* dmvrMv, size_t numDmvr);

/* ---- whole back end on the host cores: the CPU baseline / reference arm of bench.py ----
 * Runs one picture (same b200_picture work lists the GPU gets, references = 4 DPB slots) through the REFERENCE's kernels with
 * `threads` std::threads: K2 = real InterPrediction::motionCompensation per PU; K1 = the reference's DeQuant / invLfnst /
 * fastInvTrans / cpyResiClip pointers driven per TU (the TU walk is restated, the arithmetic is the reference's); K3 = real
 * LoopFilter::loopFilterCTU; K4 = real SAOProcessCTU; K5 = real ALF prepareCTU + processCTU. Reference-picture border extension
 * of one picture is included (every picture is extended once when it becomes a reference, DecLibRecon.cpp:236).
 * Returns the seconds spent in the timed part (setup of the fake vvdec objects excluded); out (may be NULL) receives the picture. */
double ref_decompress_picture_mt(const b200_geom* g, const int16_t* const* refs, const b200_picture* pic, int threads, int simd);
/* Quant::DeQuantScaling / DeQuantScalingCore (Quant.cpp:182): dequantisation with a per-position scaling-list table (piDequantCoef). */
void ref_dequant_scaling(int simd, int width, int maxX, int maxY, int scaleQP, const int32_t* dequantCoef, const int16_t* q, size_t qStride, int32_t* coef,
                         int rightShift, int inputMaximum, int32_t transformMaximum);
/* Filter flatteners (vvdec_b200/vvdec_glue/flatten_filters.h): reference structures filled from the flattened input, flattened again. */
int ref_flatten_filters(const b200_geom* g, const b200_lf_param* lfV, const b200_lf_param* lfH, const b200_sao_ctu* sao, const b200_alf_ctu* alf,
                        const b200_alf_tables* T, b200_lf_param* lfVOut, b200_lf_param* lfHOut, b200_sao_ctu* saoOut, b200_alf_ctu* alfOut,
                        int16_t* lumaCoeffOut, int16_t* lumaClipOut, int16_t* chromaCoeffOut, int16_t* chromaClipOut, int16_t* cc0Out, int16_t* cc1Out, int32_t counts[4]);
/* Flattener pin (vvdec_b200/vvdec_glue/flatten_pu.h): builds real inter CodingUnits from the syntax below, runs the real
 * InterPrediction::motionCompensation on each (prediction written to dst) and b200glue::flattenPU / flattenSbTmvp on the same CU.
 * Reference lists: L0 = {slot 0 (POC 4), slot 1 (POC 0)}, L1 = {slot 2 (POC 12), slot 3 (POC 16) or, with altRefs, slot 0 again};
 * current POC 8.  Returns the number of records written to `recs` (<= capRecs), or -1 - i if CU i was refused by the flattener. */
typedef struct ref_cu_syntax {
  int32_t x, y, w, h;
  int32_t refIdx[2];
  int32_t mv[2][3][2];            /* [list][cpmv 0..2][hor,ver], 1/16 sample */
  int32_t affine, affine6, mergeFlag, mmvdFlag, smvd, bcwIdx, imvHpel;
  int32_t sbTmvp, sbSeed;         /* MRG_TYPE_SUBPU_ATMVP with a seeded 8x8 motion field */
  int32_t geo, geoSplitDir, geoDir0, geoDir1;   /* geoFlag; interDirrefIdxGeo0/1 = (list + 1) << 4 | refIdx; partition MVs in mv[0][1], mv[1][1] */
} ref_cu_syntax;
int ref_flatten_pu_case(int simd, const b200_geom* g, const int16_t* const* refs, int altRefs, const ref_cu_syntax* cus, int numCus,
                        int16_t* const dst[3], b200_pu* recs, int capRecs, int32_t* dmvrMv, int numDmvr);
/* Explicit weighted prediction for the following ref_mc_predict / ref_decompress_picture_* calls: raw[list][refIdx][comp][3] =
 * (log2WeightDenom, iWeight, iOffset) as parsed into Slice::m_weightPredTable (pps_weighted_bipred on); NULL switches it off. */
void ref_set_wp(const int32_t* raw);
/* The application's plane writer (App/vvdecapp/vvdecHelper.h:63 _writeComponentToFile) into a memory stream: fmt 1 = pyuv, 2 = 8 bit. Returns bytes written. */
size_t ref_write_component(const int16_t* src, ptrdiff_t stride, int w, int h, int fmt, uint8_t* dst, size_t cap);
/* one intra CU of a test layout (luma samples; single tree; one TU) */
typedef struct ref_intra_cu { uint16_t x, y, w, h; uint8_t dirL, dirC, multiRefIdx, bdpcm, bdpcmC, rsv[3]; } ref_intra_cu;
/* the real IntraPrediction on the last CU of the list, or (all != 0) on every CU in order with the reconstruction step for CUs with rsv[1] set
 * (see ref_shim.cpp); rsv[0]: luma-only CU; rsv[2]: bit 0 MIP CU (dirL = MIP mode index), bit 1 transposed, bit 2 CIIP CU (its samples = the inter
 * prediction), bit 3 plain inter CU (samples given, not predicted).  Returns the number of records, < 0 on error */
int ref_intra_case(int simd, const b200_geom* g, int16_t* const planes[3], const int16_t* const resi[3], const ref_intra_cu* cus, int numCus, int all,
                   b200_intra_tu* recs, int capRecs, int cclmCollocated /* sps_chroma_vertical_collocated_flag */);
/* the real FilmGrain (updateFGC + SIMD line kernels) on the last of `frames` frames; tables / line seeds of that frame are returned (see ref_shim.cpp) */
int ref_film_grain(const int* sei, int scalarImpl, int bitDepth, int w, int h, int frames, int16_t* const planes[3], const ptrdiff_t strides[3],
                   int8_t* pattern, uint8_t* sLUT, uint8_t* pLUT, uint32_t* lineSeeds, int* scaleShift, uint8_t* compPresent);
/* calcCRC (method 1) / calcChecksum (method 2) / calcMD5 (method 0) of a 4:2:0 picture (CommonLib/PicYuvMD5.cpp); returns the digest length in bytes */
int ref_picture_hash(int method, int bitDepth, int16_t* const planes[3], const ptrdiff_t strides[3], int w, int h, uint8_t* digest, int cap);
/* LMCS through the real Reshape class (CommonLib/Reshape.cpp) and the PelBufferOps pointers it dispatches to.
 * ref_lmcs_build: createDec + the SliceReshapeInfo fields + constructReshaper(); fills `out` (tables) and invLut[1 << bitDepth]; returns 0 if the model is legal. */
int  ref_lmcs_build(int bitDepth, int minBin, int maxBin, const int* deltaCW, int chrResScalingOffset, int chromaAdj, b200_lmcs* out, int16_t* invLut);
void ref_lmcs_fwd_block(int simd, int16_t* ptr, ptrdiff_t stride, int w, int h);                 /* Reshape::rspBufFwd */
void ref_lmcs_inv_block(int simd, int16_t* ptr, ptrdiff_t stride, int w, int h);                 /* what rspCtuBcw applies to a CTU */
void ref_lmcs_scale_block(int16_t* ptr, ptrdiff_t stride, int w, int h, int scale, int bitDepth); /* AreaBuf<Pel>::scaleSignal */
int  ref_lmcs_vpdu_scale(const b200_geom* g, int16_t* const planes[3], int x, int y);            /* calculateChromaAdjVpduNei on a picture with one CU per CTU */
double ref_decompress_picture_out(const b200_geom* g, const int16_t* const* refs, const b200_picture* pic, int threads, int simd, int16_t* const out[3]);

/* ---- the DecLibRecon seam, executed (oracle/ref_seam.h): a synthetic PARSED picture (real CodingStructure: CU / TU lists through the reference's
 * allocators and Partitioner, levels in the reconstruction plane, CtuData, slice / APS state; motion left as merge / AMVP syntax) reconstructed by
 * (i) the reference's own DecLibRecon with a ThreadPool of `threads` threads, or (ii) b200glue::DecLibReconB200 (the GPU back end, same calls). ---- */
enum { SEAM_BDOF = 1, SEAM_DMVR = 2, SEAM_BCW = 4, SEAM_PROF = 8, SEAM_MMVD = 16, SEAM_GEO = 32, SEAM_CIIP = 64, SEAM_SMVD = 128, SEAM_AMVR = 256,
       SEAM_MTS = 512, SEAM_LFNST = 1024, SEAM_SBT = 2048, SEAM_MRL = 4096, SEAM_MIP = 8192, SEAM_CCLM = 16384, SEAM_JCCR = 32768, SEAM_TS = 65536,
       SEAM_BDPCM = 1 << 17, SEAM_SAO = 1 << 18, SEAM_ALF = 1 << 19, SEAM_LMCS = 1 << 20, SEAM_DEPQUANT = 1 << 21, SEAM_LOCAL_DUAL_TREE = 1 << 22,
       SEAM_VIRTUAL_BOUNDARIES = 1 << 23 /* picture header announces virtual boundaries (none are listed): a picture DecLibReconB200 refuses */,
       SEAM_NO_LF_ACROSS_SLICES = 1 << 24 /* pps_loop_filter_across_slices_enabled_flag = 0 (with numSlices > 1) */,
       SEAM_WP = 1 << 25 /* explicit weighted prediction (pps_weighted_pred / bipred), a random pred_weight_table per slice */,
       SEAM_SCALING_LIST = 1 << 26 /* explicit scaling lists: a scaling-list APS with random matrices */ };
typedef struct ref_seam_cfg {
  uint32_t seed;
  int32_t  sliceType;        /* 0 B, 1 P, 2 I                                                                  */
  int32_t  tools;            /* SEAM_* : SPS / PH / slice tool flags                                            */
  int32_t  qp;               /* slice QP; CU QPs walk around it                                                 */
  int32_t  intraPct;         /* intra CUs in P / B slices, percent                                              */
  int32_t  skipPct, mergePct, affinePct, biPct;
  int32_t  rootCbfPct, cbfPct;
  int32_t  splitPct;         /* probability of splitting a 64-sample-wide block further (scaled for other sizes) */
  int32_t  ispPct;           /* > 0 switches sps ISP on                                                         */
  int32_t  mvdSigmaQpel;     /* sigma of the MVDs in quarter samples                                            */
  int32_t  lmcsMinBin, lmcsMaxBin, lmcsDeltaCW[16], lmcsChrOffset, lmcsChromaAdj;   /* LMCS APS syntax (SEAM_LMCS) */
  int32_t  numSlices;        /* raster-scan slices of about equal size (0 / 1: one); odd slices list their reference pictures in the opposite order, use other
                              deblocking offsets and list their ALF luma APSs backwards; every third slice has CC-ALF and Cr ALF switched off        */
} ref_seam_cfg;
/* refs[slot*3+comp]: 4 reference pictures as in ref_mc_predict (unused for I pictures); filt: deblocking offsets / SAO / ALF parameters of the picture
 * (b200_picture::lfSlices, sao, alf, alfTabs and the DEBLOCK / SAO / ALF flags; the other members are ignored). */
void*  ref_seam_create(const b200_geom* g, const ref_seam_cfg* cfg, const int16_t* const* refs, const b200_picture* filt);
void   ref_seam_destroy(void* h);
size_t ref_seam_col_motion_bytes(void* h);
void   ref_seam_stats(void* h, int32_t st[16]);
/* Each handle can be reconstructed ONCE (reconstruction overwrites the levels, MIDER overwrites the motion syntax).  Both return the seconds between
 * decompressPicture() and the return of waitForPrevDecompressedPic(), < 0 on error (-4: the picture uses a tool the device path refuses). */
double ref_seam_run_stock(void* h, int threads, int16_t* const out[3], uint8_t* colMotion, size_t colBytes);
double ref_seam_run_b200(void* h, int threads, int dry, int16_t* const out[3], uint8_t* colMotion, size_t colBytes, b200_picture* flat);
double ref_seam_run_pipelined(void* const* hs, int n, int threads, int backend, int depth);   /* DecLib's alternating recon instances, see ref_seam.h */
void ref_seam_read_out(void* h, int16_t* const out[3], uint8_t* colMotion, size_t colBytes);
void* ref_seam_create_chained(const b200_geom* g, const ref_seam_cfg* cfg, const int16_t* const* refs, const b200_picture* filt, void* prev);   /* prev (POC 8) = first list-0 reference of the new picture (POC 10) */
int ref_seam_pipelined_flat(int k, b200_picture* flat);

#ifdef __cplusplus
}
#endif
#endif
