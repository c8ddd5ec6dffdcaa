// k2_inter.cu — K2: inter prediction. One CTA (256 threads) per <=16x16 luma tile of a PU (+ its two 8x8 chroma tiles):
// the reference windows of both lists are staged in shared memory (coordinates clamped to the picture = the reference's
// 144-sample border extension, Picture.cpp:400), filtered separably (8-tap luma / 4-tap chroma, 14-bit intermediates),
// then combined: rounding (uni), average / BCW (bi), BDOF, or — for DMVR tiles — a 25-point bilinear SAD search with
// parametric sub-pel refinement followed by the final MC from the padded window.  Affine PUs use a second kernel with
// one 4x4 sub-block per 16 threads (6-tap filters, PROF).  16x16 is the BDOF / DMVR processing unit of the standard,
// so tiles are independent.
//
// Replaces (reference, source/Lib/CommonLib/InterPrediction.cpp): motionCompensation :1372, xPredInterBi :686, xPredInterUni :623,
// xPredInterBlk :750, xSubPuBio :551, applyBiOptFlow :1290, BiOptFlowCore :162, gradFilterCore :212, PaddBIOCore :269,
// xProcessDMVR :1847, xinitMC :1804, xBIPMVRefine :1702, xDMVRSubPixelErrorSurface :1785, xSubPelErrorSrfc :1647,
// xPrefetchPad :1525, xFinalPaddedMCForDMVR :1731, xPredAffineBlk :934, applyPROFCore :61, xWeightedAverage :1346;
// InterpolationFilter.cpp filter<> :556, filterCopy :424, filterWxH_N4/N8 :805,:881; Buffer.cpp addAvg :441, addWeightedAvg :372;
// RdCost.cpp xGetSAD8/16(+X5) :107-221; Mv.cpp clipMvInPic :64; UnitTools.cpp PU::setAllAffineMv :2689.
// HBM traffic per bi-predicted 16x16 tile: 2*(23*23 + 2*11*11)*2 B read + (256 + 128)*2 B written + 64 B record share.
#define VVC_TABLE_QUAL static __device__ const __align__(16)
#include "vvc_tables.h"
#include "common.cuh"
#include <algorithm>
#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

namespace b200 {

constexpr int IFO = 8192;   // IF_INTERNAL_OFFS

struct McParams {
  int16_t* dst[3]; int dstStride[3];
  const int16_t* refs[B200_MAX_SLOTS * 3];   // device plane pointers, by value (no table in memory -> no sync when the DPB mapping changes)
  int refStride[3];
  int W, H, bitDepth, ctuSize, chroma;
  int fastOk;                                // plane strides are even -> rows are word-addressable
  const b200_pu* pus; const uint32_t* tiles; const int* meta;   // device lists (bucket.cu)
  int32_t* dmvrMv;
  const b200_wp* wp;                         // explicit weighted prediction entries (b200_pu::wpIdx), or null
  const b200_lmcs* lmcs; int lmcsLog2;       // LMCS: luma predictions are stored forward-mapped (DecCu.cpp:458-476); null = off
  const CUtensorMap* tmaps; uint8_t tmapBuf[B200_MAX_SLOTS];   // TMA descriptors [buffer * 3 + component] and the buffer of each slot (McLaunch)
};

// ---- TMA (cp.async.bulk.tensor): the (tw+8) x (th+7) luma and the chroma footprints of an interior 16x16 tile arrive as one bulk tensor copy each, issued
// by one thread and counted on an mbarrier; the other 63 threads go straight to the wait.  Boxes: Y 24x23, Cb / Cr 16x11 samples (rows of 48 / 32 bytes).
constexpr int TMA_LW = 24, TMA_LH = 23, TMA_CW = 16, TMA_CH = 11;
constexpr int TMA_LWIN = 576, TMA_CWIN = 192;                  // window sizes in samples, padded to 128 bytes (the destination alignment of a tensor copy)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity)
{
  asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smemDst, const CUtensorMap* map, int x, int y, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_u32(smemDst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// asynchronous 4-byte global -> shared copies (LDGSTS): a tile issues its whole footprint without waiting on any load, then waits once
__device__ __forceinline__ void cp_async4(void* smemDst, const void* gmemSrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(d), "l"(gmemSrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// final luma prediction sample -> what is stored (identity without LMCS)
#define LUMA_OUT(v) (P.lmcs ? lmcs_fwd(P.lmcs, P.lmcsLog2, (v), pmax) : (v))

struct RefPl { const int16_t* p; int w, h, stride; };
__device__ __forceinline__ int ldc(const RefPl& r, int x, int y) { return r.p[(size_t)min(max(y, 0), r.h - 1) * r.stride + min(max(x, 0), r.w - 1)]; }
// window-restricted access (DMVR prefetch padding, InterPrediction.cpp:282-318): outside [x0,x0+w) x [y0,y0+h) samples are replicas
struct Win { int x0, y0, x1, y1; };   // inclusive limits
__device__ __forceinline__ int ldw(const RefPl& r, const Win& W, int x, int y) { return ldc(r, min(max(x, W.x0), W.x1), min(max(y, W.y0), W.y1)); }

__device__ __forceinline__ void clip_mv(int& mx, int& my, int x, int y, const McParams& P)
{
  mx = clip3((-P.ctuSize - 8 - x + 1) * 16, (P.W + 8 - x - 1) * 16, mx);
  my = clip3((-P.ctuSize - 8 - y + 1) * 16, (P.H + 8 - y - 1) * 16, my);
}

__device__ __forceinline__ const int8_t* luma_taps(int frac, bool is4x4, bool altHpel)
{
  if (is4x4) return kIfLuma4x4 + frac * 8;
  if (frac == 8 && altHpel) return kIfAltHpel;
  return kIfLuma + frac * 8;
}

__device__ __forceinline__ void load_taps8(const int8_t* t, int f[8])   // rows of the tables are 8-byte aligned
{
  const int2 w = __ldg(reinterpret_cast<const int2*>(t));
#pragma unroll
  for (int k = 0; k < 4; k++) { f[k] = (w.x << (24 - 8 * k)) >> 24; f[4 + k] = (w.y << (24 - 8 * k)) >> 24; }
}

__device__ __forceinline__ int avg_bi(int p0, int p1, int w1, int hr, int pmax)
{
  int v;
  if (w1 == 4) v = (p0 + p1 + (1 << hr) + 2 * IFO) >> (hr + 1);
  else         v = (p0 * (8 - w1) + p1 * w1 + (1 << (hr + 2)) + (IFO << 3)) >> (hr + 3);
  return clip3(0, pmax, v);
}

// explicit weighted prediction (WeightPrediction.cpp:164 addWeightBi / :238 addWeightUni) on 14-bit intermediates
__device__ __forceinline__ int wp_uni(const b200_wp* e, int comp, int p, int hr, int pmax)
{
  const int s = e->shift[comp] + hr;
  return clip3(0, pmax, ((e->w0[comp] * (p + IFO) + (s > 0 ? 1 << (s - 1) : 0)) >> s) + e->offset[comp]);
}
__device__ __forceinline__ int wp_bi(const b200_wp* e, int comp, int p0, int p1, int hr, int pmax)
{
  const int s = e->shift[comp] + hr;
  return clip3(0, pmax, (e->w0[comp] * (p0 + IFO) + e->w1[comp] * (p1 + IFO) + ((1 << s) >> 1) + e->offset[comp] * (1 << (s - 1))) >> s);
}

// GEO blending weight of sample (x, y) of component scale sc in a CU of 2^l2w x 2^l2h luma samples (xWeightedGeoBlk, InterpolationFilter.cpp:1217)
__device__ __forceinline__ int geo_weight(int splitDir, int l2w, int l2h, int x, int y, int sc)
{
  const int angle = kGeoParams[splitDir * 2], mir = kGeoAngle2Mirror[angle];
  const int16_t* wo = &kGeoWeightOffset[((splitDir * 4 + (l2h - 3)) * 4 + (l2w - 3)) * 2];
  const int row = mir == 2 ? VVC_GEO_MASK_SIZE - 1 - wo[1] - (y << sc) : wo[1] + (y << sc);
  const int col = mir == 1 ? VVC_GEO_MASK_SIZE - 1 - wo[0] - (x << sc) : wo[0] + (x << sc);
  return kGeoWeights[(kGeoAngle2Mask[angle] * VVC_GEO_MASK_SIZE + row) * VVC_GEO_MASK_SIZE + col];
}
__device__ __forceinline__ int geo_blend(int wt, int p0, int p1, int hr, int pmax)
{
  const int s = hr + 3;
  return clip3(0, pmax, (wt * p0 + (8 - wt) * p1 + (1 << (s - 1)) + (IFO << 3)) >> s);
}

__device__ __forceinline__ int shift_msb(int numer, int denom) { return numer >> (31 - __clz(denom)); }   // rightShiftMSB (:92), denom > 0

__device__ int div_for_maxq7(long long N, long long D)
{
  int sign = 0, q = 0;
  if (N < 0) { sign = 1; N = -N; }
  D <<= 3;
  if (N >= D) { N -= D; q++; }
  q <<= 1; D >>= 1;
  if (N >= D) { N -= D; q++; }
  q <<= 1;
  if (N >= (D >> 1)) q++;
  return sign ? -q : q;
}

// ------------------------------------------------------------------------------------------------ translational tiles (+BDOF, DMVR)
// MODE: 0 uni, 1 bi (average / BCW), 2 bi + BDOF, 3 DMVR (+BDOF per sub-block).  blockDim.x = tw*th (32..256): one thread per luma
// sample; tw, th are powers of two (4, 8, 16), so all index arithmetic is shifts.  Dynamic shared memory, laid out per launch class.
struct TileSmem {
  int16_t *w[2], *h[2];           // luma window (stride tw+8) and H-filtered rows (stride tw) per list
  int16_t *cw[2][2], *chf[2][2];  // chroma [list][comp]: window (stride cw+4), H-filtered rows (stride cw)
  int16_t *p[2];                  // BDOF: 14-bit predictions with ring (stride 18)
};

constexpr int HS = 16, CHS = 8;   // constant row strides of the H-filtered arrays (luma / chroma): column walks use immediate offsets
constexpr int DMVR_TAIL = 1640;   // DMVR scratch behind the windows: bilinear 20x20 x2 + their 1-sample-shifted copies (1600), later the
                                  // shifted final windows (2 x 552 luma, 4 x 132 chroma)

__host__ __device__ inline int mc_smem_elems(int mode, int n, bool tma = false)   // n = tw*th; worst case over the shapes of that size
{
  if (tma) {                                                                // 16x16 tiles with TMA windows (modes 0..2): padded windows first
    const int lists = mode == 0 ? 1 : 2;
    int e = lists * (TMA_LWIN + 23 * HS + 2 * (TMA_CWIN + 11 * CHS));
    if (mode >= 2) e = max(e, 8 * n) + 2 * 324;
    return (e + 64 + 64) & ~7;                                              // + alignment slack
  }
  // window (tw+8)x(th+7): 16x16 -> 24x23; 128 -> 24x15 | 16x23; 64 -> 12x23; 32 -> 12x15.  hf: (th+7)*HS.
  const int win = n == 256 ? 552 : n == 128 ? 368 : n == 64 ? 276 : 180;
  const int hf  = (n == 256 ? 23 : n == 128 ? 23 : n == 64 ? 23 : 15) * HS;
  const int cwin = n == 256 ? 132 : n == 128 ? 88 : n == 64 ? 66 : 42;     // (cw+4)(ch+3)
  const int chf = (n == 256 ? 11 : n == 128 ? 11 : n == 64 ? 11 : 7) * CHS;
  const int lists = mode == 0 ? 1 : 2;
  int e = lists * (win + hf + 2 * (cwin + chf));
  if (mode == 3) e = max(e, 2 * 441);                                       // boundary DMVR tiles stage raw 21x21 windows here first
  if (mode >= 2) e = max(e, 8 * n);                                         // BDOF per-sample records (16 B) reuse the window area
  if (mode == 3) e += DMVR_TAIL;
  if (mode >= 2) e += 2 * 324;                                              // P0/P1 18x18
  return (e + 64) & ~7;
}

// OPT: luma outputs per thread (1: blockDim = tw*th; 4: blockDim = tw*th/4, each thread filters 4 adjacent samples so that 11 loads feed 32 MACs)
template <int MODE, int OPT, bool TMA>
__device__ __forceinline__ void mc_tile(const McParams& P, const uint32_t tile, int16_t* smem, unsigned* sSad, int* sDec, int (*sVxy)[2], unsigned long long* sBar)
{
  const int tid = threadIdx.x, nthr = blockDim.x;
  const b200_pu& pu = P.pus[tile >> 6];
  const int puW = pu.w, puH = pu.h, puX = pu.x, puY = pu.y, flags = pu.flags;
  const int tx0 = (tile & 7) * 16, ty0 = ((tile >> 3) & 7) * 16;
  const int tw = min(16, puW - tx0), th = min(16, puH - ty0);
  const int l2w = 31 - __clz(tw);
  const int bx = puX + tx0, by = puY + ty0;
  const int bd = P.bitDepth, pmax = (1 << bd) - 1, hr = max(2, 14 - bd), sh1 = 6 - hr;
  constexpr bool BI = MODE != 0;
  const bool altHpel = flags & B200_PU_ALTHPEL;
  const bool is4x4 = puW == 4 && puH == 4;
  const int l0 = BI ? 0 : (pu.refSlot[0] >= 0 ? 0 : 1);     // first (or only) list
  constexpr int NL = BI ? 2 : 1;
  const int cw = tw >> 1, ch = th >> 1, l2cw = l2w - 1;
  const int chroma = P.chroma;
  const b200_wp* we = (MODE <= 1 && P.wp && pu.wpIdx) ? P.wp + pu.wpIdx - 1 : nullptr;   // explicit weights (never with BDOF / DMVR)
  const bool geo = MODE == 1 && (flags & B200_PU_GEO);      // geometric partitioning: the two 'lists' are the two partitions' uni-predictions
  const int gl2w = 31 - __clz(puW), gl2h = 31 - __clz(puH);
  const int WS = tw + 8, CS = TMA ? TMA_CW : cw + 4;         // window strides: even, so a row is a run of 32-bit words (TMA: the box widths)

  // ---- shared memory carve-up (strides depend on the tile shape) ----
  TileSmem S;
  int16_t* tail = smem;                                      // DMVR scratch (MODE 3)
  {
    int16_t* q = smem;
    const int win = WS * (th + 7), hf = (th + 7) * HS, cwin = CS * (ch + 3), chf = (ch + 3) * CHS;
    if (TMA) {                                               // windows first, each on a 128-byte boundary (tensor copy destinations)
      q += ((128 - (smem_u32(smem) & 127)) & 127) >> 1;      // the dynamic segment starts behind the kernel's static shared variables
#pragma unroll
      for (int l = 0; l < NL; l++) { S.w[l] = q; q += TMA_LWIN; }
#pragma unroll
      for (int l = 0; l < NL; l++)
#pragma unroll
        for (int c = 0; c < 2; c++) { S.cw[l][c] = q; q += TMA_CWIN; }
#pragma unroll
      for (int l = 0; l < NL; l++) { S.h[l] = q; q += hf; }
#pragma unroll
      for (int l = 0; l < NL; l++)
#pragma unroll
        for (int c = 0; c < 2; c++) { S.chf[l][c] = q; q += chf; }
    } else {
#pragma unroll
    for (int l = 0; l < NL; l++) { S.w[l] = q; q += win; S.h[l] = q; q += hf; }
#pragma unroll
    for (int l = 0; l < NL; l++)
#pragma unroll
      for (int c = 0; c < 2; c++) { S.cw[l][c] = q; q += cwin; S.chf[l][c] = q; q += chf; }
    }
    if (MODE == 3 && q < smem + 2 * 441) q = smem + 2 * 441;
    if (MODE >= 2) { if (q < smem + 8 * tw * th) q = smem + 8 * tw * th; if (MODE == 3) { tail = q; q += DMVR_TAIL; } S.p[0] = q; S.p[1] = q + 324; }
  }

  // ---- per-list reference planes and motion ----
  const int16_t* rp[NL][3]; int mvx[NL], mvy[NL]; int tbuf[NL];
#pragma unroll
  for (int li = 0; li < NL; li++) {
    const int l = BI ? li : l0;
    const int slot = pu.refSlot[l];
    if (TMA) tbuf[li] = P.tmapBuf[slot];
#pragma unroll
    for (int c = 0; c < 3; c++) rp[li][c] = P.refs[slot * 3 + c];
    mvx[li] = pu.mv[l][0]; mvy[li] = pu.mv[l][1];
  }
  const int W = P.W, H = P.H, CWp = W >> 1, CHp = H >> 1, rs0 = P.refStride[0], rs1 = P.refStride[1];
  int ox[NL], oy[NL], ocx[NL], ocy[NL];                       // integer reference position of output (0,0), luma / chroma
  int wx0[NL][2], wy0[NL][2], wx1[NL][2], wy1[NL][2];         // DMVR padded-window limits [list][luma|chroma]
  int fmx[NL], fmy[NL];                                       // final (clipped) MVs
  bool bio = MODE == 2;

  int wofs[NL], cofs[NL];                                     // index of the footprint's first sample in each shared window row
  bool dmvrFast = false;
  if (MODE == 3) {
    // ================================================================ DMVR search (xProcessDMVR :1847)
    // Interior tiles: the 8-tap / 4-tap footprints of the INITIAL motion are copied once (word-wide); the bilinear search window
    // (xinitMC :1804, 2 integer samples around the block) is a sub-window of the luma footprint, and the padded window of the final MC
    // (xPrefetchPad :1525 / xFinalPaddedMCForDMVR :1731) is that same footprint shifted by the integer part of the refinement and
    // clamped to it.  Boundary tiles (footprint touching the picture edge, where MV clipping may act) go sample by sample.
    int16_t* B0 = tail; int16_t* B1 = tail + 400;            // bilinear buffers 20x20 (stride 20) ...
    int16_t* B0s = tail + 800; int16_t* B1s = tail + 1200;   // ... and copies shifted by one sample, so that every SAD row is word-aligned
    const int BW = tw + 4, BH = th + 4;
    const int warp = tid >> 5, lane = tid & 31, nw = max(1, nthr >> 5);
    int ix[2], iy[2], icx[2], icy[2];
    dmvrFast = P.fastOk;
#pragma unroll
    for (int li = 0; li < 2; li++) {
      ix[li] = bx + (mvx[li] >> 4); iy[li] = by + (mvy[li] >> 4); icx[li] = (bx >> 1) + (mvx[li] >> 5); icy[li] = (by >> 1) + (mvy[li] >> 5);
      dmvrFast = dmvrFast && ix[li] >= 6 && ix[li] + tw + 8 < W && iy[li] >= 6 && iy[li] + th + 8 < H;
      dmvrFast = dmvrFast && (!chroma || (icx[li] >= 4 && icx[li] + cw + 5 < CWp && icy[li] >= 3 && icy[li] + ch + 4 < CHp));
    }
    const int16_t* braw[2]; int bstr;                        // raw integer samples of the search window: (BW+1)x(BH+1) per list
    int bxF[2], byF[2];
    if (dmvrFast) {
#pragma unroll
      for (int li = 0; li < 2; li++) {
        bxF[li] = mvx[li] & 15; byF[li] = mvy[li] & 15;
        wofs[li] = (ix[li] - 3) & 1;
        {
          const int half = lane >> 4, wl = lane & 15, rw = rs0 >> 1;
          const uint32_t* src = reinterpret_cast<const uint32_t*>(rp[li][0] + (size_t)(iy[li] - 3) * rs0 + ((ix[li] - 3) & ~1)) + (size_t)(warp * 2 + half) * rw + wl;
          uint32_t* dst = reinterpret_cast<uint32_t*>(S.w[li]) + (warp * 2 + half) * (WS >> 1) + wl;
          if (wl < (WS >> 1))
#pragma unroll 4
            for (int y = warp * 2 + half; y < th + 7; y += nw * 2) { cp_async4(dst, src); src += (size_t)(nw * 2) * rw; dst += nw * 2 * (WS >> 1); }
        }
        cofs[li] = (icx[li] - 1) & 1;
        if (chroma) {
          const int sub = lane >> 3, wl = lane & 7, rwc = rs1 >> 1, npc = (ch + 6) >> 2;   // 4 rows per pass, npc passes per component
          const size_t cbase = (size_t)(icy[li] - 1) * rs1 + ((icx[li] - 1) & ~1);
          for (int q = warp; q < 2 * npc; q += nw) {
            const int c = q >= npc, y = ((c ? q - npc : q) << 2) + sub;
            if (y < ch + 3 && wl < (CS >> 1))
              cp_async4(reinterpret_cast<uint32_t*>(c ? S.cw[li][1] : S.cw[li][0]) + y * (CS >> 1) + wl, reinterpret_cast<const uint32_t*>((c ? rp[li][2] : rp[li][1]) + cbase) + (size_t)y * rwc + wl);
          }
        }
        braw[li] = S.w[li] + WS + wofs[li] + 1;
      }
      bstr = WS;
    } else {
      int16_t* R0 = smem;                                    // raw windows, stride 21
      constexpr int NB = 8;
      int X0[2], Y0[2];
#pragma unroll
      for (int li = 0; li < 2; li++) {
        int cx = mvx[li], cy = mvy[li];
        clip_mv(cx, cy, puX, puY, P);                        // xinitMC :1811: relative to the CU
        const int mx = cx - 32, my = cy - 32;
        bxF[li] = mx & 15; byF[li] = my & 15; X0[li] = min(max(puX + tx0 + (mx >> 4) + lane, 0), W - 1); Y0[li] = puY + ty0 + (my >> 4);
        braw[li] = R0 + li * 441;
      }
      bstr = 21;
      if (lane < BW + 1)
        for (int y0 = warp; y0 < BH + 1; y0 += nw * NB) {
          int16_t v[2][NB];
#pragma unroll
          for (int li = 0; li < 2; li++)
#pragma unroll
            for (int k = 0; k < NB; k++) {
              const int y = y0 + k * nw;
              if (y < BH + 1) v[li][k] = __ldg(rp[li][0] + (size_t)min(max(Y0[li] + y, 0), H - 1) * rs0 + X0[li]);
            }
#pragma unroll
          for (int li = 0; li < 2; li++)
#pragma unroll
            for (int k = 0; k < NB; k++) {
              const int y = y0 + k * nw;
              if (y < BH + 1) R0[li * 441 + y * 21 + lane] = v[li][k];
            }
        }
    }
    if (tid < 25) sSad[tid] = 0;
    cp_async_wait_all();
    __syncthreads();
    {
      // bilinear interpolation to 10 bit (filterN2_2D, InterpolationFilter.cpp:1133): one thread per (list, column) walks down the rows
      // and keeps the previous row's horizontal result.  Two-step form; equals the reference's one-step special cases for
      // xF == 0 or yF == 0 at bit depths <= 10.  Coefficients are (16 - frac, frac).
      const int s1 = 4 - (10 - bd), o1 = 1 << (s1 - 1);
      for (int i = tid; i < 2 * BW; i += nthr) {
        const int li = i >= BW, x = li ? i - BW : i;
        const int f1 = li ? bxF[1] : bxF[0], f0 = 16 - f1, g1 = li ? byF[1] : byF[0], g0 = 16 - g1;
        const int16_t* r = (li ? braw[1] : braw[0]) + x;
        int16_t* o = (li ? B1 : B0) + x; int16_t* os = (li ? B1s : B0s) + x - 1;
        int prev = (int16_t)((f0 * r[0] + f1 * r[1] + o1) >> s1);
        for (int y = 0; y < BH; y++) {
          r += bstr;
          const int cur = (int16_t)((f0 * r[0] + f1 * r[1] + o1) >> s1);
          const int16_t v = (int16_t)((g0 * prev + g1 * cur + 8) >> 4);
          o[y * 20] = v; if (x) os[y * 20] = v;
          prev = cur;
        }
      }
    }
    __syncthreads();
    {
      // SAD over every second row (RdCost.cpp:113-135): item = (position p, row pair); samples are 10-bit non-negative, so two of them
      // go through one packed max/min/subtract, and the two running halves cannot carry (8 words * 1023 < 2^16)
      const int l2hh = 30 - __clz(th);                       // log2(th / 2)
      for (int it = tid; it < (25 << l2hh); it += nthr) {
        const int p = it >> l2hh, y = (it & ((th >> 1) - 1)) * 2;
        const int u = p % 5 - 2, v = p / 5 - 2, odd = u & 1;
        const uint32_t* a = reinterpret_cast<const uint32_t*>((odd ? B0s : B0) + (2 + v + y) * 20 + 2 + u - odd);
        const uint32_t* b = reinterpret_cast<const uint32_t*>((odd ? B1s : B1) + (2 - v + y) * 20 + 2 - u - odd);
        uint32_t acc = 0;
#pragma unroll 4
        for (int x = 0; x < (tw >> 1); x++) acc += __vmaxs2(a[x], b[x]) - __vmins2(a[x], b[x]);
        atomicAdd(&sSad[p], (acc & 0xffff) + (acc >> 16));
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned minCost = sSad[12]; minCost -= minCost >> 2;  // (:1924-1925)
      int dx = 0, dy = 0;
      if (minCost >= (unsigned)(tw * th)) {
        const unsigned c12 = minCost;
        int bi_ = 12;
        for (int i = 0; i < 25; i++) { const unsigned v = i == 12 ? c12 : sSad[i]; if (v < minCost) { minCost = v; bi_ = i; } }   // xBIPMVRefine: raster order, strict <
        const int bu = bi_ % 5 - 2, bv = bi_ / 5 - 2;
        dx = bu * 16; dy = bv * 16;
        if (abs(dx) != 32 && abs(dy) != 32) {                // xDMVRSubPixelErrorSurface / xSubPelErrorSrfc
          auto at = [&](int i) -> unsigned long long { return i == 12 ? c12 : sSad[i]; };
          const unsigned long long s0 = at(bi_), sl = at(bi_ - 1), st = at(bi_ - 5), sr = at(bi_ + 1), sb = at(bi_ + 5);
          { const long long num = (long long)(sl - sr) * 16, den = (long long)(sl + sr - (s0 << 1));
            if (den != 0) dx += (sl != s0 && sr != s0) ? div_for_maxq7(num, den) : (sl == s0 ? -8 : 8); }
          { const long long num = (long long)(st - sb) * 16, den = (long long)(st + sb - (s0 << 1));
            if (den != 0) dy += (st != s0 && sb != s0) ? div_for_maxq7(num, den) : (st == s0 ? -8 : 8); }
        }
      }
      sDec[0] = dx; sDec[1] = dy; sDec[2] = (minCost < (unsigned)(2 * tw * th)) ? 0 : 1;   // bioAppliedSubblk (:1984)
      if (P.dmvrMv) {
        const int num = (ty0 >> 4) * max(1, puW >> 4) + (tx0 >> 4);
        P.dmvrMv[(pu.dmvrOff + num) * 2] = dx; P.dmvrMv[(pu.dmvrOff + num) * 2 + 1] = dy;
      }
    }
    __syncthreads();
    bio = (flags & B200_PU_BDOF) && sDec[2];
    const int dmx = sDec[0], dmy = sDec[1];
#pragma unroll
    for (int li = 0; li < 2; li++) {
      const int mrgx = mvx[li], mrgy = mvy[li];
      const int rx = clip3(-(1 << 17), (1 << 17) - 1, li ? mrgx - dmx : mrgx + dmx), ry = clip3(-(1 << 17), (1 << 17) - 1, li ? mrgy - dmy : mrgy + dmy);
      int cx = rx, cy = ry;
      clip_mv(cx, cy, bx, by, P);                            // cMvClipped, relative to the sub-block (:1749); no-op for interior tiles
      fmx[li] = cx; fmy[li] = cy;
      if (dmvrFast) {
        // shifted + clamped copies of the footprints into the (now free) bilinear scratch
        const int dIx = (rx >> 4) - (mrgx >> 4), dIy = (ry >> 4) - (mrgy >> 4);
        if (dIx | dIy) {
          const int16_t* T = S.w[li] + wofs[li]; int16_t* D = tail + li * 552;
          const int xs = min(max(lane + dIx, 0), tw + 6);
          if (lane < tw + 7)
            for (int y = warp; y < th + 7; y += nw) D[y * WS + lane] = T[min(max(y + dIy, 0), th + 6) * WS + xs];
          S.w[li] = D; wofs[li] = 0;
        }
        const int dCx = (rx >> 5) - (mrgx >> 5), dCy = (ry >> 5) - (mrgy >> 5);
        if (chroma && (dCx | dCy)) {
          const int c = lane >> 4, xl = lane & 15;
          const int16_t* T = (c ? S.cw[li][1] : S.cw[li][0]) + cofs[li]; int16_t* D = tail + 1104 + (li * 2 + c) * 132;
          const int xs = min(max(xl + dCx, 0), cw + 2);
          if (xl < cw + 3)
            for (int y = warp; y < ch + 3; y += nw) D[y * CS + xl] = T[min(max(y + dCy, 0), ch + 2) * CS + xs];
          S.cw[li][0] = tail + 1104 + (li * 2) * 132; S.cw[li][1] = tail + 1104 + (li * 2 + 1) * 132; cofs[li] = 0;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 2; k++) {                        // k = 0 luma, 1 chroma (xFinalPaddedMCForDMVR :1757-1778, xPrefetchPad :1525)
          const int sh = 4 + k, taps = k ? 4 : 8;
          const int dIx = (rx >> sh) - (mrgx >> sh), dIy = (ry >> sh) - (mrgy >> sh);
          int X, Y;
          if (dIx || dIy) {
            int pmx = mrgx - ((taps / 2 - 1) << sh), pmy = mrgy - ((taps / 2 - 1) << sh);
            clip_mv(pmx, pmy, bx, by, P);
            wx0[li][k] = (bx >> k) + (pmx >> sh); wy0[li][k] = (by >> k) + (pmy >> sh);
            wx1[li][k] = wx0[li][k] + (tw >> k) + taps - 2; wy1[li][k] = wy0[li][k] + (th >> k) + taps - 2;
            X = wx0[li][k] + (taps / 2 - 1) + dIx; Y = wy0[li][k] + (taps / 2 - 1) + dIy;
          } else {
            wx0[li][k] = wy0[li][k] = -(1 << 20); wx1[li][k] = wy1[li][k] = 1 << 20;
            X = (bx >> k) + (cx >> sh); Y = (by >> k) + (cy >> sh);
          }
          if (k == 0) { ox[li] = X; oy[li] = Y; } else { ocx[li] = X; ocy[li] = Y; }
        }
      }
    }
  } else {
#pragma unroll
    for (int li = 0; li < NL; li++) {
      int cx = mvx[li], cy = mvy[li];
      clip_mv(cx, cy, puX, puY, P);                          // relative to the CU (xPredInterUni :651)
      fmx[li] = cx; fmy[li] = cy;
      ox[li] = bx + (cx >> 4); oy[li] = by + (cy >> 4); ocx[li] = (bx >> 1) + (cx >> 5); ocy[li] = (by >> 1) + (cy >> 5);
    }
  }

  // ================================================================ stage A: windows (luma 8-tap footprint, chroma 4-tap footprint)
  // Interior tiles (the footprint lies inside the picture and no DMVR window clamp applies) copy whole 32-bit words, 16 lanes per luma
  // row / 8 lanes per chroma row; the footprint's first sample then sits at index wofs (0/1) of each shared row.  Boundary tiles take
  // the per-sample path with clamped coordinates (= the reference's border extension / padded DMVR window).
  bool tmaAny = false;
  if (!(MODE == 3 && dmvrFast)) {
    const int warp = tid >> 5, lane = tid & 31, nw = max(1, nthr >> 5);
    bool lfast[NL], cfastv[NL];
#pragma unroll
    for (int li = 0; li < NL; li++) {
      lfast[li] = P.fastOk && ox[li] >= 4 && ox[li] + tw + 4 < W && oy[li] >= 3 && oy[li] + th + 3 < H;
      cfastv[li] = chroma && P.fastOk && ocx[li] >= 2 && ocx[li] + cw + 2 < CWp && ocy[li] >= 1 && ocy[li] + ch + 1 < CHp;
      if (MODE == 3) { lfast[li] = lfast[li] && wx1[li][0] == (1 << 20); cfastv[li] = cfastv[li] && wx1[li][1] == (1 << 20); }
      tmaAny = tmaAny || lfast[li] || cfastv[li];
    }
    if (TMA && tmaAny && tid == 0) {
      // interior footprints: one tensor copy per window (the boxes may hang over the right / bottom picture edge in their padding columns only: zero fill)
      int bytes = 0;
#pragma unroll
      for (int li = 0; li < NL; li++) bytes += (lfast[li] ? TMA_LW * TMA_LH * 2 : 0) + (cfastv[li] ? 2 * TMA_CW * TMA_CH * 2 : 0);
      mbar_expect_tx(sBar, bytes);
#pragma unroll
      for (int li = 0; li < NL; li++) {
        const CUtensorMap* tm = P.tmaps + tbuf[li] * 3;
        if (lfast[li]) tma_load_2d(S.w[li], tm, ox[li] - 3, oy[li] - 3, sBar);
        if (cfastv[li]) { tma_load_2d(S.cw[li][0], tm + 1, ocx[li] - 1, ocy[li] - 1, sBar); tma_load_2d(S.cw[li][1], tm + 2, ocx[li] - 1, ocy[li] - 1, sBar); }
      }
    }
#pragma unroll
    for (int li = 0; li < NL; li++) {
      const bool fast = lfast[li];
      wofs[li] = (fast && !TMA) ? ((ox[li] - 3) & 1) : 0;
      if (fast && TMA) {
      } else if (fast) {
        const int half = lane >> 4, wl = lane & 15, rw = rs0 >> 1;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(rp[li][0] + (size_t)(oy[li] - 3) * rs0 + ((ox[li] - 3) & ~1)) + (size_t)(warp * 2 + half) * rw + wl;
        uint32_t* dst = reinterpret_cast<uint32_t*>(S.w[li]) + (warp * 2 + half) * (WS >> 1) + wl;
        if (wl < (WS >> 1))
#pragma unroll 4
          for (int y = warp * 2 + half; y < th + 7; y += nw * 2) { cp_async4(dst, src); src += (size_t)(nw * 2) * rw; dst += nw * 2 * (WS >> 1); }
      } else {
        const int xlo = MODE == 3 ? clip3(0, W - 1, wx0[li][0]) : 0, xhi = MODE == 3 ? clip3(0, W - 1, wx1[li][0]) : W - 1;
        const int ylo = MODE == 3 ? clip3(0, H - 1, wy0[li][0]) : 0, yhi = MODE == 3 ? clip3(0, H - 1, wy1[li][0]) : H - 1;
        const int xc = min(max(ox[li] - 3 + lane, xlo), xhi);
        if (lane < tw + 7)
#pragma unroll 4
          for (int y = warp; y < th + 7; y += nw) S.w[li][y * WS + lane] = __ldg(rp[li][0] + (size_t)min(max(oy[li] - 3 + y, ylo), yhi) * rs0 + xc);
      }
      cofs[li] = 0;
      if (chroma) {
        const bool cfast = cfastv[li];
        cofs[li] = (cfast && !TMA) ? ((ocx[li] - 1) & 1) : 0;
        if (cfast && TMA) {
        } else if (cfast) {
          const int sub = lane >> 3, wl = lane & 7, rwc = rs1 >> 1, npc = (ch + 6) >> 2;   // 4 rows per pass, npc passes per component
          const size_t cbase = (size_t)(ocy[li] - 1) * rs1 + ((ocx[li] - 1) & ~1);
          for (int q = warp; q < 2 * npc; q += nw) {
            const int c = q >= npc, y = ((c ? q - npc : q) << 2) + sub;
            if (y < ch + 3 && wl < ((cw + 4) >> 1))
              cp_async4(reinterpret_cast<uint32_t*>(c ? S.cw[li][1] : S.cw[li][0]) + y * (CS >> 1) + wl, reinterpret_cast<const uint32_t*>((c ? rp[li][2] : rp[li][1]) + cbase) + (size_t)y * rwc + wl);
          }
        } else {
          // both chroma components: lanes 0..15 Cb, 16..31 Cr (cw+3 <= 11)
          const int c = lane >> 4, xl = lane & 15;
          const int xlo = MODE == 3 ? clip3(0, CWp - 1, wx0[li][1]) : 0, xhi = MODE == 3 ? clip3(0, CWp - 1, wx1[li][1]) : CWp - 1;
          const int ylo = MODE == 3 ? clip3(0, CHp - 1, wy0[li][1]) : 0, yhi = MODE == 3 ? clip3(0, CHp - 1, wy1[li][1]) : CHp - 1;
          const int xcc = min(max(ocx[li] - 1 + xl, xlo), xhi);
          const int16_t* rc = c ? rp[li][2] : rp[li][1];
          int16_t* dc = c ? S.cw[li][1] : S.cw[li][0];
          if (xl < cw + 3)
#pragma unroll 4
            for (int y = warp; y < ch + 3; y += nw) dc[y * CS + xl] = __ldg(rc + (size_t)min(max(ocy[li] - 1 + y, ylo), yhi) * rs1 + xcc);
        }
      }
    }
  }
  cp_async_wait_all();
  if (TMA && tmaAny) mbar_wait(sBar, 0);
  __syncthreads();

  // ================================================================ stage B: horizontal filters
#pragma unroll
  for (int li = 0; li < NL; li++) {
    int f[8];
    load_taps8(luma_taps(fmx[li] & 15, is4x4, altHpel), f);  // row 0 of the tables is {0,0,0,64,0,0,0,0}: full-pel is the same formula
    const int16_t* sw = S.w[li] + wofs[li];
    if (OPT == 1) {
      for (int i = tid; i < (th + 7) << l2w; i += nthr) {
        const int y = i >> l2w, x = i & (tw - 1);
        const int16_t* s = sw + y * WS + x;
        int a = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) a += f[k] * s[k];
        S.h[li][y * HS + x] = (int16_t)((a - (IFO << sh1)) >> sh1);
      }
    } else {
      const int l2g = l2w - 2;                               // groups of 4 outputs per row
      for (int i = tid; i < (th + 7) << l2g; i += nthr) {
        const int y = i >> l2g, x = (i & ((tw >> 2) - 1)) << 2;
        const int16_t* s = sw + y * WS + x;
        int v[11];
#pragma unroll
        for (int k = 0; k < 11; k++) v[k] = s[k];
        int16_t* o = S.h[li] + y * HS + x;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int a = 0;
#pragma unroll
          for (int k = 0; k < 8; k++) a += f[k] * v[j + k];
          o[j] = (int16_t)((a - (IFO << sh1)) >> sh1);
        }
      }
    }
    if (chroma) {
      const int8_t* t = kIfChroma + (fmx[li] & 31) * 4;
      const int c0 = t[0], c1 = t[1], c2 = t[2], c3 = t[3];
      for (int i = tid; i < 2 * ((ch + 3) << l2cw); i += nthr) {
        const int c = i >= ((ch + 3) << l2cw), j = c ? i - ((ch + 3) << l2cw) : i;
        const int y = j >> l2cw, x = j & (cw - 1);
        const int16_t* s = (c ? S.cw[li][1] : S.cw[li][0]) + cofs[li] + y * CS + x;
        (c ? S.chf[li][1] : S.chf[li][0])[y * CHS + x] = (int16_t)((c0 * s[0] + c1 * s[1] + c2 * s[2] + c3 * s[3] - (IFO << sh1)) >> sh1);
      }
    }
  }
  __syncthreads();

  // ================================================================ stage C: vertical filters + combine
  // thread -> samples (x, y0 .. y0+OPT-1); with BDOF the same thread keeps its samples' terms in registers until the final combine
  const int sx = tid & (tw - 1), sy = (tid >> l2w) * OPT;
  {
    int pr[NL][OPT];
#pragma unroll
    for (int li = 0; li < NL; li++) {
      int f[8];
      load_taps8(luma_taps(fmy[li] & 15, is4x4, altHpel), f);
      const int16_t* s = S.h[li] + sy * HS + sx;
      int v[7 + OPT];
#pragma unroll
      for (int k = 0; k < 7 + OPT; k++) v[k] = s[k * HS];
#pragma unroll
      for (int j = 0; j < OPT; j++) {
        int a = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) a += f[k] * v[j + k];
        pr[li][j] = a;
      }
    }
#pragma unroll
    for (int j = 0; j < OPT; j++) {
      int16_t* d = P.dst[0] + (size_t)(by + sy + j) * P.dstStride[0] + bx + sx;
      if (MODE == 1 && geo) *d = (int16_t)LUMA_OUT(geo_blend(geo_weight(pu.bcwW1, gl2w, gl2h, tx0 + sx, ty0 + sy + j, 0), (int16_t)(pr[0][j] >> 6), (int16_t)(pr[NL - 1][j] >> 6), hr, pmax));
      else if (MODE <= 1 && we) *d = (int16_t)LUMA_OUT(BI ? wp_bi(we, 0, (int16_t)(pr[0][j] >> 6), (int16_t)(pr[NL - 1][j] >> 6), hr, pmax) : wp_uni(we, 0, (int16_t)(pr[0][j] >> 6), hr, pmax));
      else if (!BI) *d = (int16_t)LUMA_OUT(clip3(0, pmax, (pr[0][j] + (1 << (5 + hr)) + (IFO << 6)) >> (6 + hr)));
      else if (!bio) *d = (int16_t)LUMA_OUT(avg_bi((int16_t)(pr[0][j] >> 6), (int16_t)(pr[NL - 1][j] >> 6), MODE == 1 ? pu.bcwW1 : 4, hr, pmax));
      else { S.p[0][(sy + j + 1) * 18 + sx + 1] = (int16_t)(pr[0][j] >> 6); S.p[1][(sy + j + 1) * 18 + sx + 1] = (int16_t)(pr[NL - 1][j] >> 6); }
    }
  }
  if (chroma) {
    for (int i = tid; i < 2 * (ch << l2cw); i += nthr) {
      const int c = i >= (ch << l2cw), j = c ? i - (ch << l2cw) : i;
      const int y = j >> l2cw, x = j & (cw - 1);
      int pr[NL];
#pragma unroll
      for (int li = 0; li < NL; li++) {
        const int8_t* t = kIfChroma + (fmy[li] & 31) * 4;
        const int16_t* s = (c ? S.chf[li][1] : S.chf[li][0]) + y * CHS + x;
        pr[li] = t[0] * s[0] + t[1] * s[CHS] + t[2] * s[2 * CHS] + t[3] * s[3 * CHS];
      }
      int16_t* d = (c ? P.dst[2] : P.dst[1]) + (size_t)((by >> 1) + y) * (c ? P.dstStride[2] : P.dstStride[1]) + (bx >> 1) + x;
      if (MODE == 1 && geo) *d = (int16_t)geo_blend(geo_weight(pu.bcwW1, gl2w, gl2h, (tx0 >> 1) + x, (ty0 >> 1) + y, 1), (int16_t)(pr[0] >> 6), (int16_t)(pr[NL - 1] >> 6), hr, pmax);
      else if (MODE <= 1 && we) *d = (int16_t)(BI ? wp_bi(we, 1 + c, (int16_t)(pr[0] >> 6), (int16_t)(pr[NL - 1] >> 6), hr, pmax) : wp_uni(we, 1 + c, (int16_t)(pr[0] >> 6), hr, pmax));
      else if (!BI) *d = (int16_t)clip3(0, pmax, (pr[0] + (1 << (5 + hr)) + (IFO << 6)) >> (6 + hr));
      else *d = (int16_t)avg_bi((int16_t)(pr[0] >> 6), (int16_t)(pr[NL - 1] >> 6), MODE == 1 ? pu.bcwW1 : 4, hr, pmax);
    }
  }

  // ================================================================ BDOF (applyBiOptFlow :1290)
  if (MODE >= 2) {
    if (!bio) return;                                        // CTA-uniform
    // ring of integer reference samples around the block (xPredInterBlk :847-885); the windows already hold them
    for (int i = tid; i < 4 * (tw + th + 2); i += nthr) {
      const int li = i >= 2 * (tw + th + 2), j = li ? i - 2 * (tw + th + 2) : i;
      int x, y;
      if (j < tw + 2) { x = j; y = 0; } else if (j < 2 * (tw + 2)) { x = j - (tw + 2); y = th + 1; }
      else if (j < 2 * (tw + 2) + th) { x = 0; y = j - 2 * (tw + 2) + 1; } else { x = tw + 1; y = j - 2 * (tw + 2) - th + 1; }
      const int mxl = li ? fmx[NL - 1] : fmx[0], myl = li ? fmy[NL - 1] : fmy[0];
      const int xo = (mxl & 15) < 8 ? 1 : 0, yo = (myl & 15) < 8 ? 1 : 0;
      const int16_t* wl = li ? S.w[NL - 1] + wofs[NL - 1] : S.w[0] + wofs[0];
      const int v = wl[(y - yo + 3) * WS + (x - xo + 3)];
      (li ? S.p[1] : S.p[0])[y * 18 + x] = (int16_t)((int16_t)(v << hr) - IFO);
    }
    __syncthreads();
    // per sample: gradients of both lists (gradFilterCore<true> :212) folded into the five summands of calcBIOSums (:134), one 16-byte
    // record per sample; the padded border of the reference (PaddBIOCore :269, replication of the outermost samples) becomes a clamp of
    // the record index when the 6x6 windows are summed.  The terms of the final combine stay in this thread's registers.
    uint4* Q = reinterpret_cast<uint4*>(smem);
    int dgx[OPT], dgy[OPT], psum[OPT];
    {
      int q0[OPT + 2], q1[OPT + 2];                          // column sx, rows sy-1 .. sy+OPT (>> 6)
#pragma unroll
      for (int r = 0; r < OPT + 2; r++) { q0[r] = S.p[0][(sy + r) * 18 + sx + 1] >> 6; q1[r] = S.p[1][(sy + r) * 18 + sx + 1] >> 6; }
#pragma unroll
      for (int j = 0; j < OPT; j++) {
        const int i = (sy + j + 1) * 18 + sx + 1;
        const int p0 = S.p[0][i], p1 = S.p[1][i];
        const int g0x = (S.p[0][i + 1] >> 6) - (S.p[0][i - 1] >> 6), g1x = (S.p[1][i + 1] >> 6) - (S.p[1][i - 1] >> 6);
        const int g0y = q0[j + 2] - q0[j], g1y = q1[j + 2] - q1[j];
        const int tX = (g0x + g1x) >> 1, tY = (g0y + g1y) >> 1, dI = (p1 >> 4) - (p0 >> 4);
        uint4 rec;
        rec.x = (unsigned)abs(tX) | ((unsigned)abs(tY) << 16);
        rec.y = (unsigned)(tX < 0 ? -dI : (tX == 0 ? 0 : dI));
        rec.z = (unsigned)(tY < 0 ? -dI : (tY == 0 ? 0 : dI));
        rec.w = (unsigned)(tY < 0 ? -tX : (tY == 0 ? 0 : tX));
        Q[((sy + j) << l2w) + sx] = rec;
        dgx[j] = g0x - g1x; dgy[j] = g0y - g1y; psum[j] = p0 + p1;
      }
    }
    __syncthreads();
    const int nBlk = (tw >> 2) * (th >> 2), l2bw = l2w - 2;
    // 4 lanes per 4x4 block: lane part p sums window rows p and p+4 (rows 4,5 only for p<2); quad shuffle reduce; lane 0 derives (vx,vy)
    for (int it = tid; it < ((nBlk * 4 + 31) & ~31); it += nthr) {
      const int blk = it >> 2, part = it & 3;
      unsigned sA = 0; int sDX = 0, sDY = 0, sS = 0;
      if (blk < nBlk) {
        const int bxx = (blk & ((tw >> 2) - 1)) << 2, byy = (blk >> l2bw) << 2;
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
          const int yy = part + 4 * rr;
          if (yy < 6) {
            const uint4* row = Q + (min(max(byy + yy - 1, 0), th - 1) << l2w);
#pragma unroll
            for (int xx = 0; xx < 6; xx++) {
              const uint4 r = row[min(max(bxx + xx - 1, 0), tw - 1)];
              sA += r.x; sDX += (int)r.y; sDY += (int)r.z; sS += (int)r.w;
            }
          }
        }
      }
#pragma unroll
      for (int m = 1; m < 4; m <<= 1) {
        sA += __shfl_xor_sync(0xffffffffu, sA, m);
        sDX += __shfl_xor_sync(0xffffffffu, sDX, m); sDY += __shfl_xor_sync(0xffffffffu, sDY, m);
        sS  += __shfl_xor_sync(0xffffffffu, sS, m);
      }
      if (part == 0 && blk < nBlk) {
        const int sAX = sA & 0xffff, sAY = sA >> 16;         // 36 * 256 < 2^16: the packed halves cannot carry
        int vx = sAX == 0 ? 0 : shift_msb(sDX * 4, sAX);
        vx = clip3(-15, 15, vx);
        const int mainG = sS >> 12, secG = sS & 4095;
        int tmp = vx * mainG;
        tmp = ((tmp * (1 << 12)) + vx * secG) >> 1;
        int vy = sAY == 0 ? 0 : shift_msb(sDY * 4 - tmp, sAY);
        vy = clip3(-15, 15, vy);
        sVxy[blk][0] = vx; sVxy[blk][1] = vy;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < OPT; j++) {                          // addBIOAvg4 (:109)
      const int y = sy + j, blk = ((y >> 2) << l2bw) + (sx >> 2);
      const int b = sVxy[blk][0] * dgx[j] + sVxy[blk][1] * dgy[j];
      const int shiftNum = 15 - bd, offset = (1 << (shiftNum - 1)) + 2 * IFO;
      P.dst[0][(size_t)(by + y) * P.dstStride[0] + bx + sx] = (int16_t)LUMA_OUT(clip3(0, pmax, (int)(int16_t)((psum[j] + b + offset) >> shiftNum)));
    }
  }
}

// One CTA per entry of one tile list (bucket.cu); the host sizes the grid from the list length it reads back after bucketing, and the
// hardware scheduler balances the lists of all streams.
template <int MODE, int OPT, bool TMA>
__global__ void __launch_bounds__(64, MODE >= 2 ? 12 : 16) mc_kernel(const McParams P, const int list)
{
  extern __shared__ __align__(128) int16_t smem[];
  __shared__ unsigned sSad[25];
  __shared__ int sDec[3];
  __shared__ int sVxy[16][2];
  __shared__ __align__(8) unsigned long long sBar;
  if ((int)blockIdx.x >= P.meta[LM_CNT + list]) return;
  if (TMA) { if (threadIdx.x == 0) mbar_init(&sBar, 1); __syncthreads(); }
  mc_tile<MODE, OPT, TMA>(P, P.tiles[P.meta[LM_OFF + list] + blockIdx.x], smem, sSad, sDec, sVxy, &sBar);
}

// ------------------------------------------------------------------------------------------------ affine tiles (xPredAffineBlk :934)
__device__ __forceinline__ void round_affine(int& x, int& y, int s) { const int o = 1 << (s - 1); x = (x + o - (x >= 0)) >> s; y = (y + o - (y >= 0)) >> s; }

__device__ bool spread_over_limit(int a, int b, int c, int d, int predType)
{
  const int s4 = 4 << 11, ft = 6;
  if (predType == 3) {
    int rw = max(max(0, 4 * a + s4), max(4 * c, 4 * a + 4 * c + s4)) - min(min(0, 4 * a + s4), min(4 * c, 4 * a + 4 * c + s4));
    int rh = max(max(0, 4 * b), max(4 * d + s4, 4 * b + 4 * d + s4)) - min(min(0, 4 * b), min(4 * d + s4, 4 * b + 4 * d + s4));
    rw = (rw >> 11) + ft + 3; rh = (rh >> 11) + ft + 3;
    return rw * rh > (ft + 9) * (ft + 9);
  }
  int rw = max(0, 4 * a + s4) - min(0, 4 * a + s4), rh = max(0, 4 * b) - min(0, 4 * b);
  rw = (rw >> 11) + ft + 3; rh = (rh >> 11) + ft + 3;
  if (rw * rh > (ft + 9) * (ft + 5)) return true;
  rw = max(0, 4 * c) - min(0, 4 * c); rh = max(0, 4 * d + s4) - min(0, 4 * d + s4);
  rw = (rw >> 11) + ft + 3; rh = (rh >> 11) + ft + 3;
  return rw * rh > (ft + 5) * (ft + 9);
}

struct AffModel { int LTx, LTy, dHX, dHY, dVX, dVY; bool over, prof; };

__device__ void aff_model(const b200_pu& pu, int l, AffModel& M)
{
  const int l2w = 31 - __clz((int)pu.w), l2h = 31 - __clz((int)pu.h);
  M.LTx = pu.mv[l][0]; M.LTy = pu.mv[l][1];
  const int RTx = pu.cpmv[l][0][0], RTy = pu.cpmv[l][0][1], LBx = pu.cpmv[l][1][0], LBy = pu.cpmv[l][1][1];
  const bool six = pu.flags & B200_PU_AFFINE6;
  M.dHX = (RTx - M.LTx) * (1 << (7 - l2w)); M.dHY = (RTy - M.LTy) * (1 << (7 - l2w));
  M.dVX = six ? (LBx - M.LTx) * (1 << (7 - l2h)) : -M.dHY; M.dVY = six ? (LBy - M.LTy) * (1 << (7 - l2h)) : M.dHX;
  M.over = spread_over_limit(M.dHX, M.dHY, M.dVX, M.dVY, pu.interDir);
  bool prof = pu.flags & (l ? B200_PU_PROF1 : B200_PU_PROF0);
  if (six ? (M.LTx == RTx && M.LTy == RTy && M.LTx == LBx && M.LTy == LBy) : (M.LTx == RTx && M.LTy == RTy)) prof = false;
  if (M.over) prof = false;
  M.prof = prof;
}

// MV of luma 4x4 sub-block (i,j) of the PU (PU::setAllAffineMv, UnitTools.cpp:2689), not yet picture-clipped
__device__ __forceinline__ void aff_sub_mv(const AffModel& M, const b200_pu& pu, int i, int j, int& mx, int& my)
{
  if (M.over) { mx = M.LTx * 128 + M.dHX * (pu.w >> 1) + M.dVX * (pu.h >> 1); my = M.LTy * 128 + M.dHY * (pu.w >> 1) + M.dVY * (pu.h >> 1); }
  else        { mx = M.LTx * 128 + M.dHX * (2 + 4 * i) + M.dVX * (2 + 4 * j); my = M.LTy * 128 + M.dHY * (2 + 4 * i) + M.dVY * (2 + 4 * j); }
  round_affine(mx, my, 7);
  mx = clip3(-(1 << 17), (1 << 17) - 1, mx); my = clip3(-(1 << 17), (1 << 17) - 1, my);
}

__device__ __forceinline__ void mc_affine_tile(const McParams& P, const uint32_t tile)
{
  __shared__ int16_t sHf[2][16][9 * 4];     // per sub-block: 9 rows x 4 cols after horizontal filter
  __shared__ int16_t sE[2][16][36];         // PROF 6x6 buffers
  __shared__ int16_t sP[2][3][256];         // 14-bit predictions (luma 16x16, chroma 8x8 each)
  __shared__ int16_t sCH[2][2][4][7 * 4];   // chroma: [list][comp][sub-block] 7 rows x 4 cols

  const int tid = threadIdx.x;
  const b200_pu& pu = P.pus[tile >> 6];                      // read in place: per-list fields are indexed at run time, a private copy would live in local memory
  const int tx0 = (tile & 7) * 16, ty0 = ((tile >> 3) & 7) * 16;
  const int tw = min(16, pu.w - tx0), th = min(16, pu.h - ty0);
  const int bx = pu.x + tx0, by = pu.y + ty0;
  const int bd = P.bitDepth, pmax = (1 << bd) - 1, hr = max(2, 14 - bd), sh1 = 6 - hr;
  const bool bi = pu.refSlot[0] >= 0 && pu.refSlot[1] >= 0;
  const b200_wp* we = (P.wp && pu.wpIdx) ? P.wp + pu.wpIdx - 1 : nullptr;
  const int nList = bi ? 2 : 1, l0 = pu.refSlot[0] >= 0 ? 0 : 1;
  const int hMin = (-P.ctuSize - 8 - pu.x + 1) * 16, hMax = (P.W + 8 - pu.x - 1) * 16;
  const int vMin = (-P.ctuSize - 8 - pu.y + 1) * 16, vMax = (P.H + 8 - pu.y - 1) * 16;


  // ---- luma: sub-block sb = tid>>4 (4x4 grid in the tile), lane k = tid&15 ----
  const int sb = tid >> 4, k = tid & 15;
  const int sbx = (sb & 3) * 4, sby = (sb >> 2) * 4;
  const bool sbValid = sbx < tw && sby < th;
  for (int li = 0; li < nList; li++) {
    const int l = bi ? li : l0;
    AffModel Ml; aff_model(pu, l, Ml);                       // per list, in registers
    RefPl Rl; Rl.p = P.refs[pu.refSlot[l] * 3]; Rl.w = P.W; Rl.h = P.H; Rl.stride = P.refStride[0];
    int mx = 0, my = 0;
    if (sbValid) { aff_sub_mv(Ml, pu, (tx0 + sbx) >> 2, (ty0 + sby) >> 2, mx, my); mx = clip3(hMin, hMax, mx); my = clip3(vMin, vMax, my); }
    const int xF = mx & 15, yF = my & 15, X0 = bx + sbx + (mx >> 4), Y0 = by + sby + (my >> 4);
    if (sbValid) {
      const int8_t* fh = kIfLuma4x4 + xF * 8;
      for (int j = k; j < 36; j += 16) {                      // 9 rows (y-2..y+6: 6-tap taps 1..6 of the 8-tap array) x 4 cols
        const int y = j >> 2, x = j & 3;
        int s = 0;
        if (xF == 0) s = 64 * ldc(Rl, X0 + x, Y0 + y - 2);
        else {
#pragma unroll
          for (int t = 1; t < 7; t++) s += fh[t] * ldc(Rl, X0 + x + t - 3, Y0 + y - 2);
        }
        sHf[l][sb][j] = (int16_t)((s - (IFO << sh1)) >> sh1);
      }
      if (Ml.prof) {                                       // ring of the 6x6 PROF buffer from integer samples (:1233-1262)
        const int rx = X0 + (xF >> 3) - 1, ry = Y0 + (yF >> 3) - 1;
        for (int j = k; j < 36; j += 16) {
          const int y = j / 6, x = j - y * 6;
          if (x > 0 && x < 5 && y > 0 && y < 5) continue;
          sE[l][sb][j] = (int16_t)((int16_t)(ldc(Rl, rx + x, ry + y) << hr) - IFO);
        }
      }
    }
    __syncwarp();
    if (sbValid) {
      const int y = k >> 2, x = k & 3;
      const int8_t* fv = kIfLuma4x4 + yF * 8;
      int s = 0;
      if (yF == 0) s = 64 * sHf[l][sb][(y + 2) * 4 + x];
      else {
#pragma unroll
        for (int t = 1; t < 7; t++) s += fv[t] * sHf[l][sb][(y + t - 1) * 4 + x];
      }
      if (Ml.prof) sE[l][sb][(y + 1) * 6 + x + 1] = (int16_t)(s >> 6);
      else if (bi)   sP[l][0][(sby + y) * 16 + sbx + x] = (int16_t)(s >> 6);
      else P.dst[0][(size_t)(by + sby + y) * P.dstStride[0] + bx + sbx + x] = (int16_t)LUMA_OUT(we ? wp_uni(we, 0, (int16_t)(s >> 6), hr, pmax) : clip3(0, pmax, (s + (1 << (5 + hr)) + (IFO << 6)) >> (6 + hr)));
    }
    __syncwarp();
    if (sbValid && Ml.prof) {                              // gradFilterCore<false> :212 + applyPROFCore :61
      const int y = k >> 2, x = k & 3, c = (y + 1) * 6 + x + 1;
      const int16_t* E = sE[l][sb];
      const int gX = (E[c + 1] >> 6) - (E[c - 1] >> 6), gY = (E[c + 6] >> 6) - (E[c - 6] >> 6);
      // dMv of sample (x,y) inside the 4x4 (:1043-1090): linear in x,y, then rounded by 8 and clipped to +-31
      const int qHX = Ml.dHX * 4, qHY = Ml.dHY * 4, qVX = Ml.dVX * 4, qVY = Ml.dVY * 4;
      int dh = ((Ml.dHX + Ml.dVX) * 2) - ((qHX + qVX) * 2) + x * qHX + y * qVX;
      int dv = ((Ml.dHY + Ml.dVY) * 2) - ((qHY + qVY) * 2) + x * qHY + y * qVY;
      round_affine(dh, dv, 8);
      dh = clip3(-31, 31, dh); dv = clip3(-31, 31, dv);
      const int lim = 1 << max(bd + 1, 13);
      const int dI = clip3(-lim, lim - 1, dh * gX + dv * gY);
      int v = (int16_t)(E[c] + dI);
      if (bi) sP[l][0][(sby + y) * 16 + sbx + x] = (int16_t)v;
      else P.dst[0][(size_t)(by + sby + y) * P.dstStride[0] + bx + sbx + x] = (int16_t)LUMA_OUT(we ? wp_uni(we, 0, v, hr, pmax) : clip3(0, pmax, (int)(int16_t)((v + (1 << (hr - 1)) + IFO) >> hr)));
    }
  }

  // ---- chroma 4:2:0: 4x4 chroma sub-blocks (= 8x8 luma), MV = rounded mean of the TL and BR luma sub-block MVs (:1135-1151) ----
  if (P.chroma) {
    // jobs: (list, comp, chroma sub-block 0..3): 16 threads each -> nList*2*4*16 = 128/256 threads
    const int job = tid >> 4;
    const int li = job >> 3, c = 1 + ((job >> 2) & 1), cs = job & 3;
    const int l = bi ? li : l0;
    const int csx = (cs & 1) * 4, csy = (cs >> 1) * 4;       // chroma offset inside the 8x8 chroma tile
    const bool valid = li < nList && csx < (tw >> 1) && csy < (th >> 1);
    int mx = 0, my = 0;
    RefPl Rc; Rc.p = nullptr; Rc.w = P.W >> 1; Rc.h = P.H >> 1; Rc.stride = P.refStride[1];
    if (valid) {
      AffModel Ml; aff_model(pu, l, Ml);
      Rc.p = P.refs[pu.refSlot[l] * 3 + c]; Rc.stride = c == 1 ? P.refStride[1] : P.refStride[2];
      int ax, ay, bxm, bym;
      const int i0 = (tx0 >> 2) + (csx >> 1), j0 = (ty0 >> 2) + (csy >> 1);
      aff_sub_mv(Ml, pu, i0, j0, ax, ay); aff_sub_mv(Ml, pu, i0 + 1, j0 + 1, bxm, bym);
      mx = ax + bxm; my = ay + bym;
      round_affine(mx, my, 1);
      mx = clip3(hMin, hMax, mx); my = clip3(vMin, vMax, my);
    }
    const int xF = mx & 31, yF = my & 31, X0 = (bx >> 1) + csx + (mx >> 5), Y0 = (by >> 1) + csy + (my >> 5);
    if (valid) {
      const int8_t* fh = kIfChroma + xF * 4;
      for (int j = k; j < 28; j += 16) {
        const int y = j >> 2, x = j & 3;
        int s = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) s += fh[t] * ldc(Rc, X0 + x + t - 1, Y0 + y - 1);
        sCH[li][c - 1][cs][j] = (int16_t)((s - (IFO << sh1)) >> sh1);
      }
    }
    __syncwarp();
    if (valid) {
      const int y = k >> 2, x = k & 3;
      const int8_t* fv = kIfChroma + yF * 4;
      int s = 0;
#pragma unroll
      for (int t = 0; t < 4; t++) s += fv[t] * sCH[li][c - 1][cs][(y + t) * 4 + x];
      if (bi) sP[l][c][(csy + y) * 8 + csx + x] = (int16_t)(s >> 6);
      else P.dst[c][(size_t)((by >> 1) + csy + y) * P.dstStride[c] + (bx >> 1) + csx + x] = (int16_t)(we ? wp_uni(we, c, (int16_t)(s >> 6), hr, pmax) : clip3(0, pmax, (s + (1 << (5 + hr)) + (IFO << 6)) >> (6 + hr)));
    }
  }
  if (!bi) return;
  __syncthreads();
  {
    const int y = tid >> 4, x = tid & 15;
    if (x < tw && y < th) P.dst[0][(size_t)(by + y) * P.dstStride[0] + bx + x] = (int16_t)LUMA_OUT(we ? wp_bi(we, 0, sP[0][0][y * 16 + x], sP[1][0][y * 16 + x], hr, pmax) : avg_bi(sP[0][0][y * 16 + x], sP[1][0][y * 16 + x], pu.bcwW1, hr, pmax));
    if (P.chroma && tid < 128) {
      const int c = tid >> 6, j = tid & 63, yy = j >> 3, xx = j & 7;
      if (xx < (tw >> 1) && yy < (th >> 1))
        P.dst[1 + c][(size_t)((by >> 1) + yy) * P.dstStride[1 + c] + (bx >> 1) + xx] = (int16_t)(we ? wp_bi(we, 1 + c, sP[0][1 + c][yy * 8 + xx], sP[1][1 + c][yy * 8 + xx], hr, pmax) : avg_bi(sP[0][1 + c][yy * 8 + xx], sP[1][1 + c][yy * 8 + xx], pu.bcwW1, hr, pmax));
    }
  }
}

__global__ void __launch_bounds__(256) mc_affine_kernel(const McParams P)
{
  if ((int)blockIdx.x >= P.meta[LM_CNT + 16]) return;
  mc_affine_tile(P, P.tiles[P.meta[LM_OFF + 16] + blockIdx.x]);
}

int launch_mc(const McLaunch& L, StreamSet& ss, KProf* prof)
{
  McParams P;
  for (int c = 0; c < 3; c++) { P.dst[c] = L.dst.p[c]; P.dstStride[c] = L.dst.stride[c]; P.refStride[c] = L.refStride[c]; }
  for (int i = 0; i < B200_MAX_SLOTS * 3; i++) P.refs[i] = L.refs[i];
  P.fastOk = !(L.refStride[0] & 1) && !(L.refStride[1] & 1) && !(L.refStride[2] & 1);
  P.W = L.geom.width; P.H = L.geom.height; P.bitDepth = L.geom.bitDepth; P.ctuSize = L.geom.ctuSize; P.chroma = L.geom.chromaFormat == 1;
  P.pus = L.pus; P.dmvrMv = L.dmvrMv; P.tiles = L.tiles; P.meta = L.meta;
  P.wp = L.wp;
  P.lmcs = L.lmcs; P.lmcsLog2 = 0; { int o = (1 << L.geom.bitDepth) / 16; while ((1 << (P.lmcsLog2 + 1)) <= o) P.lmcsLog2++; }
  P.tmaps = reinterpret_cast<const CUtensorMap*>(L.tmaps); memcpy(P.tmapBuf, L.tmapBuf, sizeof(P.tmapBuf));
  // The tensor-copy windows are OFF unless B200_MC_TMA=1: on the B200 boxes of this project's pool every cp.async.bulk.tensor — this kernel's, the stand-alone
  // tools/tma_probe.cu, and libcu++'s own wrappers in tools/tma_probe_ref.cu (descriptor as __grid_constant__ parameter, in __constant__ or in global memory,
  // with and without a cluster launch) — ends in "an illegal instruction was encountered" at the UTMALDG, while the descriptor-less cp.async.bulk
  // (tools/bulk_probe.cu) and cuBLAS's own TMA kernels run.  The path is kept for a box where it runs; LDGSTS windows are the default.
  static const bool tmaEnv = getenv("B200_MC_TMA") && !strcmp(getenv("B200_MC_TMA"), "1");
  const bool tmaOn = P.tmaps && P.chroma && tmaEnv;
  int launched = 0;
  if (prof) prof->begin(B200_KF_MC_TILE, ss.main);
  for (int m = 3; m >= 0; m--) for (int k = 3; k >= 0; k--) {      // heaviest lists first (DMVR 16x16 ... uni 8x4)
    const int nsamp = 32 << k, list = m * 4 + k, grid = L.cnt[list];
    if (!grid || (m >= 2 && k < 2)) continue;                       // BDOF / DMVR tiles have at least 128 samples (bucket.cu rejects others)
    cudaStream_t s = ss.pick(launched++);
    const bool tma = tmaOn && k == 3 && m <= 2;               // 16x16 tiles of the uni / bi / BDOF lists: windows by tensor copies
    const size_t smem = (size_t)mc_smem_elems(m, nsamp, tma) * 2;
    if (tma) {
      switch (m) {
        case 0: mc_kernel<0, 4, true><<<grid, 64, smem, s>>>(P, list); break;
        case 1: mc_kernel<1, 4, true><<<grid, 64, smem, s>>>(P, list); break;
        default: mc_kernel<2, 4, true><<<grid, 64, smem, s>>>(P, list); break;
      }
    } else if (k >= 2) {     // 128 / 256 samples: 4 luma outputs per thread -> 32 / 64 threads
      const int nthr = nsamp >> 2;
      switch (m) {
        case 0: mc_kernel<0, 4, false><<<grid, nthr, smem, s>>>(P, list); break;
        case 1: mc_kernel<1, 4, false><<<grid, nthr, smem, s>>>(P, list); break;
        case 2: mc_kernel<2, 4, false><<<grid, nthr, smem, s>>>(P, list); break;
        default: mc_kernel<3, 4, false><<<grid, nthr, smem, s>>>(P, list); break;
      }
    } else {
      switch (m) {
        case 0: mc_kernel<0, 1, false><<<grid, nsamp, smem, s>>>(P, list); break;
        default: mc_kernel<1, 1, false><<<grid, nsamp, smem, s>>>(P, list); break;
      }
    }
    B200_CUDA(cudaGetLastError());
  }
  if (L.cnt[16]) { cudaStream_t s = ss.pick(launched++); mc_affine_kernel<<<L.cnt[16], 256, 0, s>>>(P); B200_CUDA(cudaGetLastError()); }
  ss.join();
  if (prof) prof->end(B200_KF_MC_TILE, ss.main);      // with forked streams the affine tiles are inside the same interval
  return 0;
}

// Tensor maps of the picture buffers: 2-D, 16-bit elements, no swizzle / interleave, zero fill outside (never read: TMA windows are interior footprints).
int make_mc_tensor_maps(const b200_geom& g, int16_t* const* bufPlanes, int nBufs, void** out)
{
  *out = nullptr;
  if (g.chromaFormat != 1 || (g.stride[0] & 7) || (g.stride[1] & 7) || (g.stride[2] & 7)) return 0;      // global strides must be multiples of 16 bytes
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) { cudaGetLastError(); return 0; }
  std::vector<CUtensorMap> maps((size_t)nBufs * 3);
  for (int i = 0; i < nBufs * 3; i++) {
    const int c = i % 3;
    if ((uintptr_t)bufPlanes[i] & 15) return 0;
    const cuuint64_t dims[2] = {(cuuint64_t)(c ? g.width >> 1 : g.width), (cuuint64_t)(c ? g.height >> 1 : g.height)};
    const cuuint64_t strides[1] = {(cuuint64_t)g.stride[c] * 2};
    const cuuint32_t box[2] = {(cuuint32_t)(c ? TMA_CW : TMA_LW), (cuuint32_t)(c ? TMA_CH : TMA_LH)}, estr[2] = {1, 1};
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(&maps[i], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, bufPlanes[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 0;
  }
  B200_CUDA(cudaMalloc(out, maps.size() * sizeof(CUtensorMap)));
  B200_CUDA(cudaMemcpy(*out, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  return 0;
}

int mc_launch_count(const McLaunch& L) { int n = L.cnt[16] > 0; for (int m = 0; m < 4; m++) for (int k = 0; k < 4; k++) n += L.cnt[m * 4 + k] > 0 && !(m >= 2 && k < 2); return n; }

}  // namespace b200
