// k6_intra.cu — K6: regular intra prediction (planar, DC, angular incl. wide angles, multi reference line, PDPC, BDPCM prediction) of a
// list of transform blocks in decoding order, with the residual add that makes a block's reconstruction the next block's reference.
// Replaces (reference, source/Lib/CommonLib/IntraPrediction.cpp): xFillReferenceSamples :1072 (sample copies / substitution),
// xFilterReferenceSamples :1251, predIntraAng :474, xPredIntraPlanarCore :154, xGetPredValDc :412, xPredIntraAng :592,
// IntraPredAngleCore :301, IntraPredAngleChroma :333, IntraPredSampleFilterCore :212, xPredIntraBDPCM :850 and the pred + resi clip of
// DecCu::predAndReco (DecCu.cpp:390-398).
//
// Scheduling.  Intra blocks depend on their neighbours' reconstruction, so the list is a dataflow graph.  CTAs take blocks in order from
// a ticket counter; a block waits (bounded spin on per-block `done` words) for the earlier blocks of the list that own the units its
// available reference samples lie in (`owner` maps filled by a pre-pass, one word per 4x4 luma / 2x2 chroma unit).  Tickets are handed out
// in decoding order and only to running CTAs, so the oldest unfinished block never waits on a block that has not started: no deadlock,
// whatever the residency.  Reference samples are read with ld.global.cg (L2): another SM wrote them.
// Inside a block every sample is independent once the two reference arrays are in shared memory: one thread computes several samples.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>
#define VVC_TABLE_QUAL static __device__ const __align__(16)
#include "vvc_tables.h"

namespace b200 {

constexpr int IT_THREADS = 128;
constexpr int IT_REF = 2 * 64 + 8;           // top / left array: 2*size + 1 + multiRefIdx entries
constexpr int IT_ORG = 72;                   // origin of the main / side arrays (room for the negative extension, <= 64)
constexpr int IT_ARR = IT_ORG + 2 * 64 + 80; // main / side arrays incl. replication tail ((mrl << s) + 2 <= 34)

__constant__ int cAng[32] = {0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29, 32, 35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512, 1024};
__constant__ int cInvAng[32] = {0, 16384, 8192, 5461, 4096, 2731, 2048, 1638, 1365, 1170, 1024, 910, 819, 712, 630, 565,
                                512, 468, 420, 364, 321, 287, 256, 224, 191, 161, 128, 96, 64, 48, 32, 16};
__constant__ int cIntraFilterThr[8] = {24, 24, 24, 14, 2, 0, 0, 0};

struct IntraParams {
  int16_t* planes[3]; const int16_t* resi[3]; int stride[3]; int W, H, bitDepth;
  const b200_intra_tu* tus; int numTus;
  int* owner[3]; int ownerStride[3];        // per unit: index of the list entry that writes it, -1: not written by this list
  int* done; int* ticket; int* err;
  const int* perm;                          // processing order (wavefront over CTUs), or null: list order
  int* ctuCnt; int* ctuFirst; int* ctuBase; int ctuLog2, ctusW, ctusH;
  int compSel;                              // 0 all blocks, 1 luma only, 2 chroma only (second launch on the same list: the luma blocks' done words are set)
};

__device__ __forceinline__ int wide_angle(int w, int h, int mode)
{
  if (mode > 1 && mode <= 66) {
    const int shift[6] = {0, 6, 10, 12, 14, 15};
    const int d = abs((31 - __clz(w)) - (31 - __clz(h)));
    if (w > h && mode < 2 + shift[d]) mode += 65;
    else if (h > w && mode > 66 - shift[d]) mode -= 65;
  }
  return mode;
}

__global__ void __launch_bounds__(256) intra_owner_kernel(const IntraParams P)
{
  const int i = blockIdx.x;
  const b200_intra_tu t = P.tus[i];
  const int c = t.comp, unit = c ? 2 : 4, uw = max(1, (1 << t.log2w) / unit), uh = max(1, (1 << t.log2h) / unit);
  // ISP regions can be 1 or 2 rows high: several share a unit, the last of them completes it
  for (int k = threadIdx.x; k < uw * uh; k += blockDim.x) atomicMax(&P.owner[c][(t.y / unit + k / uw) * P.ownerStride[c] + t.x / unit + k % uw], i);
}

// Processing order.  The list is in decoding order (CTU raster); handing tickets out in that order keeps only the next few CTUs of a CTU row
// in flight.  Any topological order of the dependency graph keeps the no-deadlock argument, and so does the wavefront order
// key(CTU) = x + 2 y (left, above, above-left and above-right CTUs all have smaller keys): CTUs of one anti-diagonal run side by side.
// perm = blocks sorted by key, decoding order kept inside a CTU (its blocks are contiguous in the list).
__device__ __forceinline__ int intra_ctu_of(const IntraParams& P, const b200_intra_tu& t)
{
  const int sh = t.comp ? 1 : 0;
  return (((int)t.y << sh) >> P.ctuLog2) * P.ctusW + (((int)t.x << sh) >> P.ctuLog2);
}
__global__ void __launch_bounds__(256) intra_ctu_count_kernel(const IntraParams P)
{
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P.numTus) return;
  const int c = intra_ctu_of(P, P.tus[i]);
  atomicAdd(&P.ctuCnt[c], 1); atomicMin(&P.ctuFirst[c], i);
}
__global__ void intra_ctu_base_kernel(const IntraParams P)
{
  int run = 0;
  for (int d = 0; d < P.ctusW + 2 * P.ctusH; d++)
    for (int y = 0; y < P.ctusH; y++) { const int x = d - 2 * y; if (x >= 0 && x < P.ctusW) { P.ctuBase[y * P.ctusW + x] = run; run += P.ctuCnt[y * P.ctusW + x]; } }
}
__global__ void __launch_bounds__(256) intra_perm_kernel(const IntraParams P, int* perm)
{
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P.numTus) return;
  const int c = intra_ctu_of(P, P.tus[i]), r = i - P.ctuFirst[c];
  if (r >= P.ctuCnt[c]) { atomicOr(P.err, 2); return; }       // the CTU's blocks are not contiguous in the list
  perm[P.ctuBase[c] + r] = i;
}

__device__ __forceinline__ int pix(const int16_t* __restrict__ plane, int stride, int x, int y) { return __ldcg(plane + (size_t)y * stride + x); }

// ISP regions (B200_INTRA_ISP): the neighbourhood in the record is the CU's (bx, by, bw, bh rebuilt from the region and its index); region k > 0 takes the row
// above (horizontal split) / the column left (vertical split) from the region before it, padded with its last sample, and the other side from the CU's
// reference arrays at the region's offset — or, where the CU has no neighbour on that side, the sample next to the region's corner
// (initIntraPatternChTypeISP, IntraPrediction.cpp:966-1070).  topLen / sideLen: m_topRefLength / m_leftRefLength.
#define ISP_GEOMETRY \
    const bool isp = t.flags & B200_INTRA_ISP; \
    int bx = x0, by = y0, bw = w, bh = h, ispK = 0, ispSplit = 0, resiMask = 0xff, l2tu = 16; \
    if (isp) { \
      ispSplit = t.mip & 3; ispK = (t.mip >> 2) & 3; const int nReg_ = 1 << ((t.mip >> 4) & 3); \
      if (ispSplit == 1) { bh = h * nReg_; by = y0 - ispK * h; } else { bw = w * nReg_; bx = x0 - ispK * w; } \
      int tuW_ = w; if (ispSplit == 2 && (bw == 4 || (bw == 8 && bh > 4))) tuW_ = max(bw >> 2, bh < 16 ? 16 / bh : 1);      /* transform units narrower than the region */ \
      l2tu = 31 - __clz(tuW_); resiMask = t.ciip; \
    } \
    const int topLen = isp ? bw + w : 2 * w, sideLen = isp ? bh + h : 2 * h;

#define ISP_REFERENCE_LAMBDAS \
    auto refT = [&](int j) -> int {                              /* row above (bx, by) incl. the corner: xFillReferenceSamples :1072 */ \
      if (n == 0) return 1 << (P.bitDepth - 1); \
      if (n == totalUnits) return PIXR(bx - 1 - mrl + j, by - 1 - mrl); \
      if (j <= mrl) {                                            /* corner part of the row */ \
        if (numLeft > 0) return availTL ? PIXR(bx - 1 - mrl + j, by - 1 - mrl) : PIXR(bx - 1 - mrl, by); \
        return PIXR(bx, by - 1 - mrl); \
      } \
      if (numAbove) return PIXR(bx + min(j - 1 - mrl, aboveLen - 1), by - 1 - mrl); \
      return availTL ? PIXR(bx - 1, by - 1 - mrl) : PIXR(bx - 1 - mrl, by);      /* = T[mrl]; numLeft > 0 here */ \
    }; \
    auto refL = [&](int i) -> int {                              /* left column, i >= 1 */ \
      if (n == 0) return 1 << (P.bitDepth - 1); \
      if (n == totalUnits) return PIXR(bx - 1 - mrl, by - 1 - mrl + i); \
      if (numLeft > 0) { \
        if (i <= mrl) return availTL ? PIXR(bx - 1 - mrl, by - 1 - mrl + i) : PIXR(bx - 1 - mrl, by); \
        return PIXR(bx - 1 - mrl, by + min(i - 1 - mrl, leftLen - 1)); \
      } \
      return PIXR(bx, by - 1 - mrl); \
    }; \
    auto regT = [&](int j) -> int { \
      if (!ispK) return refT(j); \
      if (ispSplit == 1) return j == 0 ? (t.lmLeft ? refL(y0 - by) : PIXR(x0, y0 - 1)) : PIXR(x0 + min(j - 1, w - 1), y0 - 1); \
      return t.lmAbove ? refT(x0 - bx + j) : PIXR(x0 - 1, y0); \
    }; \
    auto regL = [&](int i) -> int { \
      if (!ispK) return refL(i); \
      if (ispSplit == 1) return t.lmLeft ? refL(y0 - by + i) : PIXR(x0, y0 - 1); \
      return PIXR(x0 - 1, y0 + min(i - 1, h - 1)); \
    };

__global__ void __launch_bounds__(IT_THREADS, 12) intra_kernel(const IntraParams P)
{
  __shared__ int16_t sT[2][IT_REF], sL[2][IT_REF];          // [0] unfiltered, [1] filtered
  __shared__ int16_t sM[IT_ARR], sS[IT_ARR];
  __shared__ int sTicket, sSum;
  __shared__ int16_t sLm[32 * 32], sLmTop[64], sLmLeft[64];   // CCLM: down-sampled luma of the block, of the row above, of the column left
  __shared__ int sLmPar[3];
  const int tid = threadIdx.x;
  for (;;) {
    __syncthreads();
    if (tid == 0) { sTicket = atomicAdd(P.ticket, 1); sSum = 0; }
    __syncthreads();
    if (sTicket >= P.numTus) return;
    if (P.perm && (*(volatile int*)P.err & 2)) return;          // the list is not in decoding order (CTU runs not contiguous): perm holds holes, nothing is run
    const int me = P.perm ? P.perm[sTicket] : sTicket;
    const b200_intra_tu t = P.tus[me];
    if ((P.compSel == 1 && t.comp != 0) || (P.compSel == 2 && t.comp == 0)) continue;      // the other channel's pass
    const int c = t.comp, w = 1 << t.log2w, h = 1 << t.log2h, mrl = c ? 0 : t.multiRefIdx, unit = c ? 2 : 4;
    const int x0 = t.x, y0 = t.y, ps = P.stride[c], pmax = (1 << P.bitDepth) - 1;
    const int16_t* plane = P.planes[c];
    const int availTL = (t.flags & B200_INTRA_AVAIL_TL) ? 1 : 0, numAbove = t.numAbove, numLeft = t.numLeft;
    ISP_GEOMETRY

    // ---- wait for the earlier blocks this one reads from
    {
      if (ispK && tid == 0) {                                          // ISP: the region before this one is the record before it
        const volatile int* d = P.done + me - 1; const volatile int* e = P.err; int spins = 0;
        while (*d == 0) { __nanosleep(64); if (*e || ++spins > (1 << 22)) { atomicOr(P.err, 1); break; } }
      }
      for (int dep = tid; dep < 96; dep += IT_THREADS) {               // dependency slots: 0 corner, 1..32 above units, 64..95 left units
      int ux = -1, uy = -1;
      if (dep == 0) { if (availTL) { ux = bx - 1; uy = by - 1; } }
      else if (dep <= numAbove) { ux = bx + (dep - 1) * unit; uy = by - 1; }
      else if (dep - 64 >= 0 && dep - 64 < numLeft) { ux = bx - 1; uy = by + (dep - 64) * unit; }
      if (ux >= 0 && uy >= 0) {
        const int o = P.owner[c][(uy / unit) * P.ownerStride[c] + ux / unit];
        if (o >= 0 && o < me) {
          const volatile int* d = P.done + o;
          int spins = 0;
          const volatile int* e = P.err;
          while (*d == 0) { __nanosleep(64); if (*e || ++spins > (1 << 22)) { atomicOr(P.err, 1); break; } }   // bounded: a broken list must not hang the GPU
        }
      }
      }
      if (t.mode >= B200_INTRA_LM) {
        // CCLM reads reconstructed luma: the co-located block, up to 3 rows above and 3 columns left of it, as far as the templates go
        const bool aCu = t.flags & B200_INTRA_LM_ABOVE, lCu = t.flags & B200_INTRA_LM_LEFT;
        const int nA = max(w, aCu ? (t.mode == B200_INTRA_MDLM_T ? 2 * t.lmAbove : w) : 0), nL = max(h, lCu ? (t.mode == B200_INTRA_MDLM_L ? 2 * t.lmLeft : h) : 0);
        const int ux0 = max(0, (2 * x0 - (lCu ? 4 : 0)) >> 2), ux1 = min(P.W - 1, 2 * x0 + 2 * nA - 1) >> 2;
        const int uy0 = max(0, (2 * y0 - (aCu ? 4 : 0)) >> 2), uy1 = min(P.H - 1, 2 * y0 + 2 * nL - 1) >> 2;
        const int uw = ux1 - ux0 + 1, nU = uw * (uy1 - uy0 + 1);
        for (int u = tid; u < nU; u += IT_THREADS) {
          const int o = P.owner[0][(uy0 + u / uw) * P.ownerStride[0] + ux0 + u % uw];
          if (o >= 0 && o < me) {
            const volatile int* d = P.done + o; const volatile int* e = P.err;
            int spins = 0;
            while (*d == 0) { __nanosleep(64); if (*e || ++spins > (1 << 22)) { atomicOr(P.err, 1); break; } }
          }
        }
      }
      __threadfence();
    }
    __syncthreads();

    // ---- reference samples (xFillReferenceSamples): T[j] = row above incl. the corner, L[i] = left column, T[0] = L[0] = corner
    // (bx, by, bw, bh): the block the neighbourhood was analysed for — the block itself, or the CU of an ISP region
    const int predSize = 2 * bw, predHSize = 2 * bh;
    const int totalUnits = (predSize + unit - 1) / unit + (predHSize + unit - 1) / unit + 1, n = availTL + numAbove + numLeft;
    const int aboveLen = min(numAbove * unit, predSize), leftLen = min(numLeft * unit, predHSize);
    int16_t *T = sT[0], *L = sL[0];
#define PIXR(x, y) pix(plane, ps, (x), (y))
    ISP_REFERENCE_LAMBDAS
    for (int j = tid; j <= topLen + mrl; j += IT_THREADS) T[j] = (int16_t)regT(j);
    for (int i = tid + 1; i <= sideLen + mrl; i += IT_THREADS) L[i] = (int16_t)regL(i);       // L[0] is T[0], set below
#undef PIXR
    __syncthreads();
    if (tid == 0) L[0] = T[0];
    __syncthreads();
    if ((t.flags & B200_INTRA_FILTER_REF) && !c && !mrl) {     // xFilterReferenceSamples
      int16_t *FT = sT[1], *FL = sL[1];
      for (int j = tid; j <= predSize; j += IT_THREADS)
        FT[j] = j == 0 ? (int16_t)((L[1] + 2 * T[0] + T[1] + 2) >> 2) : j == predSize ? T[j] : (int16_t)((T[j + 1] + 2 * T[j] + T[j - 1] + 2) >> 2);
      for (int i = tid + 1; i <= predHSize; i += IT_THREADS)
        FL[i] = i == predHSize ? L[i] : (int16_t)((L[i + 1] + 2 * L[i] + L[i - 1] + 2) >> 2);
      __syncthreads();
      if (tid == 0) FL[0] = FT[0];
      T = FT; L = FL;
      __syncthreads();
    }

    const int mode = t.mode;
    const bool doPDPC = w >= 4 && h >= 4 && mrl == 0;
    int16_t* dst = P.planes[c] + (size_t)y0 * ps + x0;
    const int16_t* rs = (P.resi[c] && (t.flags & B200_INTRA_ADD_RESI)) ? P.resi[c] + (size_t)y0 * ps + x0 : nullptr;
    const int ciipW = isp ? 0 : t.ciip;                        // CIIP: the block holds the inter prediction; blend (predBlendIntraCiip :925-938)
#define IT_STORE(x, y, v) do { int v_ = (v); int16_t* d_ = dst + (size_t)(y) * ps + (x); if (ciipW) v_ = ((4 - ciipW) * (int)*d_ + ciipW * v_ + 2) >> 2; \
                               if (rs && ((resiMask >> ((x) >> l2tu)) & 1)) v_ = clip3(0, pmax, v_ + rs[(size_t)(y) * ps + (x)]); *d_ = (int16_t)v_; } while (0)

    if (mode == B200_INTRA_PLANAR || mode == B200_INTRA_DC) {
      int dc = 0;
      if (mode == B200_INTRA_DC) {                             // xGetPredValDc
        int part = 0;
        for (int i = tid; i < w + h; i += IT_THREADS) {
          if (i < w) { if (w >= h) part += T[mrl + 1 + i]; }
          else if (w <= h) part += L[mrl + 1 + i - w];
        }
        for (int o = 16; o; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
        if ((tid & 31) == 0) atomicAdd(&sSum, part);
        __syncthreads();
        const int denom = w == h ? w << 1 : max(w, h);
        dc = (sSum + (denom >> 1)) >> (31 - __clz(denom));
      }
      const int l2w = t.log2w, l2h = t.log2h, scale = (l2w - 2 + l2h - 2 + 2) >> 2;
      const int bl = L[h + 1], tr = T[w + 1];
      for (int k = tid; k < w * h; k += IT_THREADS) {
        const int y = k >> l2w, x = k & (w - 1);
        int v;
        if (mode == B200_INTRA_PLANAR) {                       // xPredIntraPlanarCore in closed form
          const int hor = (L[y + 1] << l2w) + (x + 1) * (tr - L[y + 1]), vert = (T[x + 1] << l2h) + (y + 1) * (bl - T[x + 1]);
          v = ((hor << l2h) + (vert << l2w) + (1 << (l2w + l2h))) >> (1 + l2w + l2h);
        } else v = dc;
        v = (int16_t)v;
        if (doPDPC) {                                          // IntraPredSampleFilterCore
          const int wT = 32 >> min(31, (y << 1) >> scale), wL = 32 >> min(31, (x << 1) >> scale);
          v = (int16_t)(v + ((wL * (L[y + 1] - v) + wT * (T[x + 1] - v) + 32) >> 6));
        }
        IT_STORE(x, y, v);
      }
    } else if (mode >= B200_INTRA_LM) {
      // ---- cross-component linear model (xGetLumaRecPixels, xGetLMParameters, predIntraChromaLM), 4:2:0
      const bool aCu = t.flags & B200_INTRA_LM_ABOVE, lCu = t.flags & B200_INTRA_LM_LEFT, colloc = t.flags & B200_INTRA_LM_COLLOCATED;
      const int ls = P.stride[0];
      const int16_t* rec = P.planes[0] + (size_t)(2 * y0) * ls + 2 * x0;
      const bool firstRowOfCtu = ((2 * y0) & ((1 << P.ctuLog2) - 1)) == 0;
      const int nTop = aCu ? (mode == B200_INTRA_MDLM_T ? 2 * t.lmAbove : w) : 0, nLeft = lCu ? (mode == B200_INTRA_MDLM_L ? 2 * t.lmLeft : h) : 0;
#define LUMA(p_, o_) ((int)__ldcg((p_) + (o_)))
      for (int k = tid; k < nTop + nLeft + w * h; k += IT_THREADS) {
        if (k < nTop) {                                          // row above the block
          const int i = k, m = (i == 0 && !lCu) ? 0 : 1;
          int v;
          if (firstRowOfCtu) { const int16_t* p = rec - ls; v = (LUMA(p, 2 * i) * 2 + LUMA(p, 2 * i - m) + LUMA(p, 2 * i + 1) + 2) >> 2; }
          else if (colloc) { const int16_t* p = rec - 2 * ls; v = (LUMA(p, 2 * i - ls) + LUMA(p, 2 * i) * 4 + LUMA(p, 2 * i - m) + LUMA(p, 2 * i + 1) + LUMA(p, 2 * i + ls) + 4) >> 3; }
          else { const int16_t* p = rec - 2 * ls; v = (LUMA(p, 2 * i) * 2 + LUMA(p, 2 * i - m) + LUMA(p, 2 * i + 1) + LUMA(p, 2 * i + ls) * 2 + LUMA(p, 2 * i - m + ls) + LUMA(p, 2 * i + 1 + ls) + 4) >> 3; }
          sLmTop[i] = (int16_t)v;
        } else if (k < nTop + nLeft) {                           // column left of the block
          const int j = k - nTop; const int16_t* p = rec - 3 + (size_t)(2 * j) * ls;
          int v;
          if (colloc) v = (LUMA(p, 1 - ((j == 0 && !aCu) ? 0 : ls)) + LUMA(p, 1) * 4 + LUMA(p, 0) + LUMA(p, 2) + LUMA(p, 1 + ls) + 4) >> 3;
          else v = (LUMA(p, 1) * 2 + LUMA(p, 0) + LUMA(p, 2) + LUMA(p, 1 + ls) * 2 + LUMA(p, ls) + LUMA(p, 2 + ls) + 4) >> 3;
          sLmLeft[j] = (int16_t)v;
        } else {                                                 // the block
          const int q = k - nTop - nLeft, j = q >> t.log2w, i = q & (w - 1), m = (i == 0 && !lCu) ? 0 : 1;
          const int16_t* p = rec + (size_t)(2 * j) * ls;
          int v;
          if (colloc) { const int up = (j == 0 && !aCu) ? 0 : ls; v = (LUMA(p, 2 * i - up) + LUMA(p, 2 * i) * 4 + LUMA(p, 2 * i - m) + LUMA(p, 2 * i + 1) + LUMA(p, 2 * i + ls) + 4) >> 3; }
          else v = (LUMA(p, 2 * i) * 2 + LUMA(p, 2 * i + 1) + LUMA(p, 2 * i - m) + LUMA(p, 2 * i + ls) * 2 + LUMA(p, 2 * i + 1 + ls) + LUMA(p, 2 * i - m + ls) + 4) >> 3;
          sLm[q] = (int16_t)v;
        }
      }
#undef LUMA
      __syncthreads();
      if (tid == 0) {                                            // xGetLMParameters: four template positions -> a, b, shift
        const int tuWU = w >> 1, tuHU = h >> 1;
        bool aboveAvail = false, leftAvail = false; int topNum = 0, leftNum = 0;
        if (mode == B200_INTRA_MDLM_T) { aboveAvail = t.lmAbove >= tuWU; topNum = 2 * t.lmAbove; }
        else if (mode == B200_INTRA_MDLM_L) { leftAvail = t.lmLeft >= tuHU; leftNum = 2 * t.lmLeft; }
        else { aboveAvail = aCu; leftAvail = lCu; topNum = w; leftNum = h; }
        const int aboveIs4 = leftAvail ? 0 : 1, leftIs4 = aboveAvail ? 0 : 1;
        const int start0 = topNum >> (2 + aboveIs4), step0 = max(1, topNum >> (1 + aboveIs4)), start1 = leftNum >> (2 + leftIs4), step1 = max(1, leftNum >> (1 + leftIs4));
        int sl[4] = {0, 0, 0, 0}, sc[4] = {0, 0, 0, 0}, cntT = 0, cntL = 0;
        if (aboveAvail) { cntT = min(topNum, (1 + aboveIs4) << 1); for (int k = 0, pos = start0; k < cntT; k++, pos += step0) { sl[k] = sLmTop[pos]; sc[k] = T[1 + pos]; } }
        if (leftAvail) { cntL = min(leftNum, (1 + leftIs4) << 1); for (int k = 0, pos = start1; k < cntL; k++, pos += step1) { sl[k + cntT] = sLmLeft[pos]; sc[k + cntT] = L[1 + pos]; } }
        if (cntT + cntL == 2) { sl[3] = sl[0]; sc[3] = sc[0]; sl[2] = sl[1]; sc[2] = sc[1]; sl[0] = sl[1]; sc[0] = sc[1]; sl[1] = sl[3]; sc[1] = sc[3]; }
        int mn0 = 0, mn1 = 2, mx0 = 1, mx1 = 3, tt;
        if (sl[mn0] > sl[mn1]) { tt = mn0; mn0 = mn1; mn1 = tt; }
        if (sl[mx0] > sl[mx1]) { tt = mx0; mx0 = mx1; mx1 = tt; }
        if (sl[mn0] > sl[mx1]) { tt = mn0; mn0 = mx0; mx0 = tt; tt = mn1; mn1 = mx1; mx1 = tt; }     // the two groups change roles
        if (sl[mn1] > sl[mx0]) { tt = mn1; mn1 = mx0; mx0 = tt; }
        const int minL = (sl[mn0] + sl[mn1] + 1) >> 1, minC = (sc[mn0] + sc[mn1] + 1) >> 1, maxL = (sl[mx0] + sl[mx1] + 1) >> 1, maxC = (sc[mx0] + sc[mx1] + 1) >> 1;
        int a = 0, b = 1 << (P.bitDepth - 1), shift = 0;
        if (leftAvail || aboveAvail) {
          const int diff = maxL - minL;
          if (diff > 0) {
            const int diffC = maxC - minC;
            int x = 31 - __clz(diff);
            const int normDiff = (diff << 4 >> x) & 15;
            const int v = (int)((0x0765544332211110ull >> (4 * (15 - normDiff))) & 15) | 8;   // DivSigTable {0,7,6,5,5,4,4,3,3,2,2,1,1,1,1,0}
            x += normDiff != 0;
            const int y = diffC == 0 ? 0 : (31 - __clz(abs(diffC))) + 1, add = 1 << y >> 1;
            a = (diffC * v + add) >> y; shift = 3 + x - y;
            if (shift < 1) { shift = 1; a = a == 0 ? 0 : a < 0 ? -15 : 15; }
            b = minC - ((a * minL) >> shift);
          } else { a = 0; b = minC; shift = 0; }
        }
        sLmPar[0] = a; sLmPar[1] = b; sLmPar[2] = shift;
      }
      __syncthreads();
      const int a = sLmPar[0], b = sLmPar[1], shift = sLmPar[2];
      for (int k = tid; k < w * h; k += IT_THREADS) { const int y = k >> t.log2w, x = k & (w - 1); IT_STORE(x, y, clip3(0, pmax, ((a * sLm[k]) >> shift) + b)); }
    } else if (mode == B200_INTRA_MIP) {
      // ---- matrix intra prediction (PredictorMIP): reduced boundary -> matrix stage -> linear up-sampling; sM holds the reduced prediction
      const int sizeId = (w == 4 && h == 4) ? 0 : (w == 4 || h == 4 || (w == 8 && h == 8)) ? 1 : 2;
      const int bdry = sizeId == 0 ? 2 : 4, red = sizeId < 2 ? 4 : 8, upH = w / red, upV = h / red, inSize = 2 * bdry;
      const int modeIdx = t.mip & 0x7f; const bool transpose = t.mip >> 7;
      int16_t* in = sS;                                          // the (rebased) input vector, [inSize]; in[8] = the offset taken out
      if (tid < inSize) {                                        // boundaryDownsampling1D of the top / left boundary, in the order the matrix wants
        const bool fromLeft = (tid >= bdry) != transpose; const int d = tid % bdry, len = fromLeft ? h : w;
        const int16_t* full = (fromLeft ? L : T) + 1;
        int v;
        if (bdry < len) { const int f = len / bdry; int sum = 0; for (int j = 0; j < f; j++) sum += full[d * f + j]; v = (sum + (f >> 1)) >> (31 - __clz(f)); }
        else v = full[d];
        in[tid] = (int16_t)v;
      }
      __syncthreads();
      const int inputOffset = in[0];
      __syncthreads();
      if (tid < inSize) in[tid] = (int16_t)(tid == 0 ? (sizeId < 2 ? (1 << (P.bitDepth - 1)) - inputOffset : 0) : in[tid] - inputOffset);
      __syncthreads();
      if (tid < red * red) {                                     // computeReducedPred: one output per thread
        const int redSize = sizeId == 2, stride = inSize - redSize;
        const uint8_t* wgt = (sizeId == 0 ? kMip4x4 + modeIdx * 64 : sizeId == 1 ? kMip8x8 + modeIdx * 128 : kMip16x16 + modeIdx * 448) + tid * stride;
        int sum = 0, acc = 0;
        for (int i = 0; i < inSize; i++) sum += in[i];
        for (int i = redSize; i < inSize; i++) acc += in[i] * wgt[i - redSize];
        const int v = clip3(0, pmax, ((acc + 32 - 32 * sum) >> 6) + inputOffset);
        sM[transpose ? (tid % red) * red + tid / red : tid] = (int16_t)v;
      }
      __syncthreads();
      const int l2H = 31 - __clz(upH), l2V = 31 - __clz(upV);
      for (int k = tid; k < w * h; k += IT_THREADS) {
        const int y = k >> t.log2w, x = k & (w - 1), kr = y / upV, i = y % upV;
        int hv[2];                                               // horizontally up-sampled rows kr - 1 (or the top boundary) and kr at column x
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int kk = kr - 1 + q;
          if (kk < 0) hv[q] = T[x + 1];
          else if (upH == 1) hv[q] = sM[kk * red + x];
          else {
            const int j = x / upH, ii = x % upH, before = j == 0 ? L[(kk + 1) * upV] : sM[kk * red + j - 1], behind = sM[kk * red + j];
            hv[q] = (int16_t)(before * upH + (upH >> 1) + (ii + 1) * (behind - before)) >> l2H;   // the reference accumulates in Pel (int16): wraps at 12 bit x 16
          }
        }
        IT_STORE(x, y, upV == 1 ? hv[1] : (int16_t)(hv[0] * upV + (upV >> 1) + (i + 1) * (hv[1] - hv[0])) >> l2V);
      }
    } else if (mode >= B200_INTRA_BDPCM_HOR) {
      for (int k = tid; k < w * h; k += IT_THREADS) { const int y = k >> t.log2w, x = k & (w - 1); IT_STORE(x, y, mode == B200_INTRA_BDPCM_HOR ? L[y + 1] : T[x + 1]); }
    } else {
      // ---- angular (xPredIntraAng): main / side reference arrays, then every sample on its own
      const int predMode = wide_angle(bw, bh, mode);                 // ISP: the CU decides (:610)
      const bool ver = predMode >= 34;
      const int angMode = ver ? predMode - 50 : -(predMode - 18), absMode = abs(angMode);
      const int invAngle = cInvAng[absMode], absAng = cAng[absMode], angle = angMode < 0 ? -absAng : absAng;
      const int16_t *mainSrc = ver ? T : L, *sideSrc = ver ? L : T;
      const int mw = ver ? w : h, mh = ver ? h : w;            // block size along the main / the side reference
      int16_t *M = sM + IT_ORG, *S = sS + IT_ORG;
      if (angle < 0) {
        for (int k = tid - mh; k <= mw + 1 + mrl; k += IT_THREADS) M[k] = k >= 0 ? mainSrc[k] : sideSrc[min((-k * invAngle + 256) >> 9, mh)];
        for (int k = tid; k <= mh + 1 + mrl; k += IT_THREADS) S[k] = sideSrc[k];
      } else {
        const int l2r = (31 - __clz(mw)) - (31 - __clz(mh)), s = max(0, l2r), maxIndex = (mrl << s) + 2, refLength = ver ? topLen : sideLen;
        for (int k = tid; k <= refLength + mrl + maxIndex; k += IT_THREADS) M[k] = mainSrc[min(k, refLength + mrl)];
        for (int k = tid; k <= (ver ? sideLen : topLen) + mrl; k += IT_THREADS) S[k] = sideSrc[k];
      }
      __syncthreads();
      const int16_t *Mp = M + mrl, *Sp = S + mrl;
      const int l2mw = 31 - __clz(mw), l2mh = 31 - __clz(mh);
      const int topLeft = T[0];
      const int scale0 = (l2mw - 2 + l2mh - 2 + 2) >> 2;
      const int lev = min(3 << scale0, mw);                    // lev[scale] = min(3, 6, 12, 24; width)
      const bool frac = (absAng & 31) != 0;
      const int diff = min(abs(predMode - 18), abs(predMode - 50));
      const bool cubic = isp || !(diff > cIntraFilterThr[(l2mw + l2mh) >> 1]) || mrl > 0;
      int angularScale = -1;
      if (angle > 0 && doPDPC) angularScale = min(2, l2mh - ((31 - __clz(3 * invAngle - 2)) - 8));
      for (int k = tid; k < mw * mh; k += IT_THREADS) {
        const int yy = k >> l2mw, xx = k & (mw - 1);
        int v;
        if (angle == 0) {
          if (doPDPC && xx < lev) { const int wL = 32 >> min(31, (xx << 1) >> scale0); v = clip3(0, pmax, (wL * (Sp[yy + 1] - topLeft) + Mp[xx + 1] * 64 + 32) >> 6); }
          else v = Mp[xx + 1];
        } else {
          const int deltaPos = angle * (1 + mrl + yy), dI = deltaPos >> 5, dF = deltaPos & 31;
          if (!frac) v = Mp[dI + 1 + xx];
          else if (c) v = (int16_t)(((32 - dF) * Mp[dI + 1 + xx] + dF * Mp[dI + 2 + xx] + 16) >> 5);
          else {
            const int16_t* p = Mp + dI + xx;
            int f0, f1, f2, f3;
            if (cubic) { f0 = kIfChroma[dF * 4]; f1 = kIfChroma[dF * 4 + 1]; f2 = kIfChroma[dF * 4 + 2]; f3 = kIfChroma[dF * 4 + 3]; }
            else { f0 = 16 - (dF >> 1); f1 = 32 - (dF >> 1); f2 = 16 + (dF >> 1); f3 = dF >> 1; }
            v = (int16_t)((f0 * p[0] + f1 * p[1] + f2 * p[2] + f3 * p[3] + 32) >> 6);
            if (cubic) v = clip3(0, pmax, v);
          }
          if (angularScale >= 0 && xx < min(3 << angularScale, mw)) {
            const int invAngleSum = 256 + (xx + 1) * invAngle, wL = 32 >> (2 * xx >> angularScale), left = Sp[yy + (invAngleSum >> 9) + 1];
            v = (int16_t)(v + ((wL * (left - v) + 32) >> 6));
          }
        }
        if (ver) IT_STORE(xx, yy, v); else IT_STORE(yy, xx, v);
      }
    }
#undef IT_STORE
    // ---- publish
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(P.done + me, 1);
  }
}


// ================================================================================================ K6 v2: one CTA per CTU, the CTU resident in shared memory
// The dependency chains of an intra picture run block to block; v1 pays a global-memory round trip per hop (ticket, owner lookup, done flag, reference
// samples through L2: ~4 us).  v2 keeps the hops on chip: a CTA owns one CTU at a time (CTUs handed out in wave-front order, key x + 2 y), loads the
// CTU's samples (inter prediction + residual reconstruction so far) and residual planes into shared memory, and its warps take the CTU's blocks in
// decoding order — one warp per block, every sample of a block independent once the reference arrays are built.  A block waits on shared-memory done
// flags for earlier blocks of its own CTU and on the global done words only for blocks of the neighbouring CTUs (left, above-left, above, above-right:
// all earlier in the wave front, so the oldest unfinished block never waits on a block that has not been started — no deadlock at any residency).
// Samples are written through: to the shared tile for the CTU's own later blocks, to the plane for the other CTUs and the in-loop filters.
// a block is worked on by a group of V2_GROUP threads, V2_GROUPS blocks of the CTU at a time (template parameters of the kernel: the chains of an intra CTU
// are Y -> Y -> Y, Cb -> Cb, Cr -> Cr, so about three blocks can run at once and the latency of one block is what counts)
constexpr int V2_LS = 136, V2_CS = 72;                       // tile row pitch in samples: a multiple of 16 bytes (16-byte asynchronous copies), rows 4 banks apart
constexpr int V2_TILE = 128 * V2_LS + 2 * 64 * V2_CS;        // samples of one tile set (Y, Cb, Cr)
constexpr int V2_RECS = 1024;                                // records staged in shared memory (a CTU with more blocks reads the rest from global memory)
constexpr int V2_FLAGS = 128 * 128 / 16 + 2 * (64 * 64 / 4); // most blocks a CTU can hold
struct V2Scratch { int16_t T[2][IT_REF], L[2][IT_REF], M[IT_ARR], S[IT_ARR], Lm[32 * 32], LmTop[64], LmLeft[64]; int LmPar[4]; int ticket, sum; };
constexpr int V2_OWN = 3 * 32 * 32;                         // owner words of the CTU's units: luma 32 x 32 (4x4 units), Cb / Cr 32 x 32 each (2x2 units)
constexpr size_t v2_smem(int groups) { return (size_t)2 * V2_TILE * sizeof(int16_t) + groups * sizeof(V2Scratch) + V2_RECS * sizeof(b200_intra_tu) + V2_OWN * sizeof(int) + V2_FLAGS + 64; }
__device__ __forceinline__ void v2_cp4(void* smemDst, const void* gmemSrc)      // asynchronous 4-byte global -> shared copy (LDGSTS): the whole CTU in flight before one wait
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(d), "l"(gmemSrc));
}
__device__ __forceinline__ void v2_cp16(void* smemDst, const void* gmemSrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(d), "l"(gmemSrc));
}
__device__ __forceinline__ void v2_cp_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

struct V2Tile {
  int16_t* rec[3]; int16_t* res[3]; int ox[3], oy[3], tw[3], th[3], ts[3];
  const int16_t* plane[3]; int ps[3];
  __device__ __forceinline__ bool inside(int c, int x, int y) const { return (unsigned)(x - ox[c]) < (unsigned)tw[c] && (unsigned)(y - oy[c]) < (unsigned)th[c]; }
  __device__ __forceinline__ int pix(int c, int x, int y) const
  {
    if (inside(c, x, y)) return rec[c][(y - oy[c]) * ts[c] + (x - ox[c])];
    return __ldcg(plane[c] + (size_t)y * ps[c] + x);
  }
};

// CTUs that hold blocks, sorted by the wave-front key (x + 2 y, then y): every CTU counts the non-empty CTUs that precede it
__global__ void __launch_bounds__(1024) intra_ctu_order_kernel(const IntraParams P, int* ctuOrder, int* counters)
{
  extern __shared__ int sCnt[];
  const int n = P.ctusW * P.ctusH;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sCnt[i] = P.ctuCnt[i];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (sCnt[i] <= 0) continue;
    const int y = i / P.ctusW, x = i - y * P.ctusW, key = x + 2 * y;
    int rank = 0;
    for (int j = 0; j < n; j++) { const int yj = j / P.ctusW, kj = (j - yj * P.ctusW) + 2 * yj; rank += (sCnt[j] > 0) && (kj < key || (kj == key && yj < y)); }
    ctuOrder[rank] = i;
  }
  if (threadIdx.x == 0) { int m = 0; for (int j = 0; j < n; j++) m += sCnt[j] > 0; counters[0] = m; counters[1] = 0; }
}
// contiguity of every CTU's blocks in the list (decoding order): entry i belongs to the run [ctuFirst, ctuFirst + ctuCnt) of its CTU
__global__ void __launch_bounds__(256) intra_ctu_check_kernel(const IntraParams P)
{
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P.numTus) return;
  const int c = intra_ctu_of(P, P.tus[i]);
  if (i - P.ctuFirst[c] >= P.ctuCnt[c]) atomicOr(P.err, 2);
}

// returns true if the block waited for lives in another CTU (its samples are then read from the plane: a device-scope fence is due)
__device__ __forceinline__ bool v2_wait(const IntraParams& P, const volatile uint8_t* sflag, int o, int me, int first)
{
  if (o < 0 || o >= me) return false;
  int spins = 0;
  if (o >= first) { while (sflag[o - first] == 0) { if (++spins > 64) __nanosleep(20); if (spins > (1 << 24)) { atomicOr(P.err, 1); break; } } return false; }
  const volatile int* d = P.done + o; const volatile int* e = P.err;
  while (*d == 0) { __nanosleep(32); if (*e || ++spins > (1 << 22)) { atomicOr(P.err, 1); break; } }
  return true;
}

#ifdef B200_K6_PROF
__device__ unsigned long long gK6Prof[48];
#define K6P(i, t0) do { if (lane == 0) atomicAdd(&gK6Prof[i], (unsigned long long)(clock64() - (t0))); } while (0)
#define K6C(i) do { if (lane == 0) atomicAdd(&gK6Prof[i], 1ull); } while (0)
#else
#define K6P(i, t0) do {} while (0)
#define K6C(i) do {} while (0)
#endif
template <int V2_GROUP, int V2_GROUPS>
__global__ void __launch_bounds__(V2_GROUP * V2_GROUPS, 1) intra_ctu_kernel(const IntraParams P, const int* __restrict__ ctuOrder, int* counters)
{
  constexpr int V2_THREADS = V2_GROUP * V2_GROUPS;
  extern __shared__ __align__(16) unsigned char smem[];
  int16_t* tileRec = reinterpret_cast<int16_t*>(smem);
  int16_t* tileRes = tileRec + V2_TILE;
  V2Scratch* scratch = reinterpret_cast<V2Scratch*>(tileRes + V2_TILE);
  b200_intra_tu* srec = reinterpret_cast<b200_intra_tu*>(scratch + V2_GROUPS);
  int* sown = reinterpret_cast<int*>(srec + V2_RECS);
  volatile uint8_t* sflag = reinterpret_cast<volatile uint8_t*>(sown + V2_OWN);
  __shared__ int sCtu, sNext;
  const int tid = threadIdx.x, lane = tid & (V2_GROUP - 1), grp = tid / V2_GROUP;      // lane: index inside the group
  const int nComp = P.planes[1] ? 3 : 1;
  const int ctuSize = 1 << P.ctuLog2;
  for (;;) {
    __syncthreads();
    if (tid == 0) { sCtu = atomicAdd(&counters[1], 1); sNext = 0; }
    __syncthreads();
    if (sCtu >= counters[0] || (*(volatile int*)P.err & 2)) return;          // bit 2: a CTU's blocks are not contiguous in the list (intra_ctu_check_kernel)
    const int ctu = ctuOrder[sCtu], first = P.ctuFirst[ctu], cnt = P.ctuCnt[ctu];
    long long tp = clock64(); (void)tp;
    V2Tile TL;
    {
      const int cx = (ctu % P.ctusW) << P.ctuLog2, cy = (ctu / P.ctusW) << P.ctuLog2;
      int16_t* r = tileRec; int16_t* q = tileRes;
      for (int c = 0; c < 3; c++) {
        const int sh = c ? 1 : 0;
        TL.ox[c] = cx >> sh; TL.oy[c] = cy >> sh; TL.tw[c] = min(ctuSize, P.W - cx) >> sh; TL.th[c] = min(ctuSize, P.H - cy) >> sh; TL.ts[c] = c ? V2_CS : V2_LS;
        TL.rec[c] = r; TL.res[c] = q; TL.plane[c] = P.planes[c]; TL.ps[c] = P.stride[c];
        r += (c ? 64 * V2_CS : 128 * V2_LS); q += (c ? 64 * V2_CS : 128 * V2_LS);
        if (c >= nComp) { TL.tw[c] = TL.th[c] = 0; }
      }
    }
    // ---- the CTU's samples, residuals and owner words -> shared memory (asynchronous 32-bit copies: the pitch is not a multiple of 16 bytes), records, flags
    for (int c = 0; c < nComp; c++) {
      const int16_t* src = P.planes[c] + (size_t)TL.oy[c] * P.stride[c] + TL.ox[c];
      const int16_t* rsrc = P.resi[c] ? P.resi[c] + (size_t)TL.oy[c] * P.stride[c] + TL.ox[c] : nullptr;
      if (!(P.stride[c] & 7) && !(TL.tw[c] & 7) && !((uintptr_t)src & 15) && !((uintptr_t)rsrc & 15)) {       // rows start on 16-byte boundaries: 8 samples per copy
        const int wv = TL.tw[c] >> 3, n = wv * TL.th[c];
        for (int k = tid; k < n; k += V2_THREADS) {
          const int y = k / wv, x = (k - y * wv) * 8;
          v2_cp16(TL.rec[c] + y * TL.ts[c] + x, src + (size_t)y * P.stride[c] + x);
          if (rsrc) v2_cp16(TL.res[c] + y * TL.ts[c] + x, rsrc + (size_t)y * P.stride[c] + x);
        }
      } else {
        const int wWords = TL.tw[c] >> 1, n = wWords * TL.th[c];
        for (int k = tid; k < n; k += V2_THREADS) {
          const int y = k / wWords, x = (k - y * wWords) * 2;
          v2_cp4(TL.rec[c] + y * TL.ts[c] + x, src + (size_t)y * P.stride[c] + x);
          if (rsrc) v2_cp4(TL.res[c] + y * TL.ts[c] + x, rsrc + (size_t)y * P.stride[c] + x);
        }
      }
      const int unit = c ? 2 : 4, uw = TL.tw[c] / unit, uh = TL.th[c] / unit;
      const int* osrc = P.owner[c] + (size_t)(TL.oy[c] / unit) * P.ownerStride[c] + TL.ox[c] / unit;
      for (int k = tid; k < uw * uh; k += V2_THREADS) { const int y = k / uw, x = k - y * uw; v2_cp4(sown + c * 1024 + y * 32 + x, osrc + (size_t)y * P.ownerStride[c] + x); }
    }
    for (int k = tid; k < min(cnt, V2_RECS) * 4; k += V2_THREADS) v2_cp4(reinterpret_cast<uint32_t*>(srec) + k, reinterpret_cast<const uint32_t*>(P.tus + first) + k);
    for (int k = tid; k < cnt; k += V2_THREADS) sflag[k] = P.compSel == 2 ? (uint8_t)(P.tus[first + k].comp == 0) : 0;      // chroma pass: the luma blocks are finished (the pass before)
    v2_cp_wait();
    __syncthreads();
    K6P(0, tp); K6C(8);       // [0] CTU set-up cycles, [8] CTU count (per group)

    // ---- the CTU's blocks, one group each, in decoding order
    V2Scratch& SC = scratch[grp];
#define V2_SYNC() do { if (V2_GROUP == 32) __syncwarp(); else asm volatile("bar.sync %0, %1;" :: "r"(grp + 1), "r"(V2_GROUP) : "memory"); } while (0)
    for (;;) {
      V2_SYNC();
      if (lane == 0) { SC.ticket = atomicAdd(&sNext, 1); SC.sum = 0; }
      V2_SYNC();
      const int k = SC.ticket;
      if (k >= cnt) break;
      tp = clock64(); const long long tb0 = tp; (void)tb0;
      const int me = first + k;
      const b200_intra_tu t = k < V2_RECS ? srec[k] : P.tus[me];
      if ((P.compSel == 1 && t.comp != 0) || (P.compSel == 2 && t.comp == 0)) continue;      // the other channel's pass
      const int c = t.comp, w = 1 << t.log2w, h = 1 << t.log2h, mrl = c ? 0 : t.multiRefIdx, unit = c ? 2 : 4;
      const int x0 = t.x, y0 = t.y, ps = P.stride[c], pmax = (1 << P.bitDepth) - 1;
      const int availTL = (t.flags & B200_INTRA_AVAIL_TL) ? 1 : 0, numAbove = t.numAbove, numLeft = t.numLeft;
      ISP_GEOMETRY
      // the tile of the block's component, in registers (indexing the per-component arrays with a run-time index would go through local memory)
      const int tox = TL.ox[c], toy = TL.oy[c], ttw = TL.tw[c], tth = TL.th[c], tts = TL.ts[c];
      const int16_t* trec = TL.rec[c]; const int16_t* tplane = TL.plane[c];
      auto pixC = [&](int x, int y) -> int {
        return ((unsigned)(x - tox) < (unsigned)ttw && (unsigned)(y - toy) < (unsigned)tth) ? (int)trec[(y - toy) * tts + (x - tox)] : (int)__ldcg(tplane + (size_t)y * ps + x);
      };
      // ---- wait for the earlier blocks this one reads from
      bool far = false;
      if (ispK && lane == 0) far |= v2_wait(P, sflag, me - 1, me, first);     // ISP: the region before this one is the record before it
      for (int dep = lane; dep < 96; dep += V2_GROUP) {                        // dependency slots: 0 corner, 1..32 above units, 64..95 left units
        int ux = -1, uy = -1;
        if (dep == 0) { if (availTL) { ux = bx - 1; uy = by - 1; } }
        else if (dep <= numAbove) { ux = bx + (dep - 1) * unit; uy = by - 1; }
        else if (dep - 64 >= 0 && dep - 64 < numLeft) { ux = bx - 1; uy = by + (dep - 64) * unit; }
        if (ux >= 0 && uy >= 0) {
          const int ush = c ? 1 : 2, ox = (ux >> ush) - (tox >> ush), oy = (uy >> ush) - (toy >> ush);          // inside the CTU: the staged owner word
          const int o = ((unsigned)ox < 32u && (unsigned)oy < 32u && ux < tox + ttw && uy < toy + tth) ? sown[c * 1024 + oy * 32 + ox] : __ldcg(P.owner[c] + (size_t)(uy >> ush) * P.ownerStride[c] + (ux >> ush));
          far |= v2_wait(P, sflag, o, me, first);
        }
      }
      if (t.mode >= B200_INTRA_LM) {
        const bool aCu = t.flags & B200_INTRA_LM_ABOVE, lCu = t.flags & B200_INTRA_LM_LEFT;
        const int nA = max(w, aCu ? (t.mode == B200_INTRA_MDLM_T ? 2 * t.lmAbove : w) : 0), nL = max(h, lCu ? (t.mode == B200_INTRA_MDLM_L ? 2 * t.lmLeft : h) : 0);
        const int ux0 = max(0, (2 * x0 - (lCu ? 4 : 0)) >> 2), ux1 = min(P.W - 1, 2 * x0 + 2 * nA - 1) >> 2;
        const int uy0 = max(0, (2 * y0 - (aCu ? 4 : 0)) >> 2), uy1 = min(P.H - 1, 2 * y0 + 2 * nL - 1) >> 2;
        const int uw = ux1 - ux0 + 1, nU = uw * (uy1 - uy0 + 1);
        for (int u = lane; u < nU; u += V2_GROUP) {
          const int gx = ux0 + u % uw, gy = uy0 + u / uw, ox = gx - TL.ox[0] / 4, oy = gy - TL.oy[0] / 4;
          const int o = ((unsigned)ox < 32u && (unsigned)oy < 32u && gx * 4 < TL.ox[0] + TL.tw[0] && gy * 4 < TL.oy[0] + TL.th[0]) ? sown[oy * 32 + ox] : __ldcg(P.owner[0] + (size_t)gy * P.ownerStride[0] + gx);
          far |= v2_wait(P, sflag, o, me, first);
        }
      }
      if (far) __threadfence(); else __threadfence_block();      // a thread that saw another CTU's done word orders the plane reads of its group behind it
      V2_SYNC();
      const int pb = c ? 16 : 0; (void)pb;
      K6P(pb + 1, tp); tp = clock64();                          // [1] dependency wait

      // ---- reference samples (xFillReferenceSamples): T[j] = row above incl. the corner, L[i] = left column, T[0] = L[0] = corner
      const int predSize = 2 * bw, predHSize = 2 * bh;
      const int totalUnits = (predSize + unit - 1) / unit + (predHSize + unit - 1) / unit + 1, n = availTL + numAbove + numLeft;
      const int aboveLen = min(numAbove * unit, predSize), leftLen = min(numLeft * unit, predHSize);
      int16_t *T = SC.T[0], *L = SC.L[0];
#define PIXR(x, y) pixC((x), (y))
      ISP_REFERENCE_LAMBDAS
      // one pass over both arrays: a thread takes entry j of the row and entry j of the column (two independent fetches in flight); entry 0 of the column is the corner
      for (int j = lane; j <= max(topLen, sideLen) + mrl; j += V2_GROUP) {
        const bool doT = j <= topLen + mrl, doL = j >= 1 && j <= sideLen + mrl;
        const int vt = doT ? regT(j) : 0, vl = doL ? regL(j) : 0;
        if (doT) T[j] = (int16_t)vt;
        if (doL) L[j] = (int16_t)vl;
        if (j == 0) L[0] = (int16_t)vt;
      }
#undef PIXR
      V2_SYNC();
      if ((t.flags & B200_INTRA_FILTER_REF) && !c && !mrl) {     // xFilterReferenceSamples
        int16_t *FT = SC.T[1], *FL = SC.L[1];
        for (int j = lane; j <= predSize; j += V2_GROUP)
          FT[j] = j == 0 ? (int16_t)((L[1] + 2 * T[0] + T[1] + 2) >> 2) : j == predSize ? T[j] : (int16_t)((T[j + 1] + 2 * T[j] + T[j - 1] + 2) >> 2);
        for (int i = lane + 1; i <= predHSize; i += V2_GROUP)
          FL[i] = i == predHSize ? L[i] : (int16_t)((L[i + 1] + 2 * L[i] + L[i - 1] + 2) >> 2);
        V2_SYNC();
        if (lane == 0) FL[0] = FT[0];
        T = FT; L = FL;
        V2_SYNC();
      }

      K6P(pb + 2, tp); tp = clock64();                          // [2] reference samples (+ filter)
      const int mode = t.mode;
      const bool doPDPC = w >= 4 && h >= 4 && mrl == 0;
      int16_t* dstG = P.planes[c] + (size_t)y0 * ps + x0;
      int16_t* dstT = const_cast<int16_t*>(trec) + (y0 - toy) * tts + (x0 - tox);
      const int tsC = tts;
      const int16_t* rsT = (P.resi[c] && (t.flags & B200_INTRA_ADD_RESI)) ? TL.res[c] + (y0 - toy) * tsC + (x0 - tox) : nullptr;
      const int ciipW = isp ? 0 : t.ciip;
#define V2_STORE(x, y, v) do { int v_ = (v); int16_t* d_ = dstT + (y) * tsC + (x); if (ciipW) v_ = ((4 - ciipW) * (int)*d_ + ciipW * v_ + 2) >> 2; \
                               if (rsT && ((resiMask >> ((x) >> l2tu)) & 1)) v_ = clip3(0, pmax, v_ + rsT[(y) * tsC + (x)]); *d_ = (int16_t)v_; dstG[(size_t)(y) * ps + (x)] = (int16_t)v_; } while (0)

      if (mode == B200_INTRA_PLANAR || mode == B200_INTRA_DC) {
        int dc = 0;
        if (mode == B200_INTRA_DC) {                             // xGetPredValDc
          int part = 0;
          for (int i = lane; i < w + h; i += V2_GROUP) {
            if (i < w) { if (w >= h) part += T[mrl + 1 + i]; }
            else if (w <= h) part += L[mrl + 1 + i - w];
          }
          for (int o = 16; o; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
          if ((lane & 31) == 0 && part) atomicAdd(&SC.sum, part);
          V2_SYNC();
          const int denom = w == h ? w << 1 : max(w, h);
          dc = (SC.sum + (denom >> 1)) >> (31 - __clz(denom));
        }
        const int l2w = t.log2w, l2h = t.log2h, scale = (l2w - 2 + l2h - 2 + 2) >> 2;
        const int bl = L[h + 1], tr = T[w + 1];
#pragma unroll 2
        for (int kk = lane; kk < w * h; kk += V2_GROUP) {
          const int y = kk >> l2w, x = kk & (w - 1);
          int v;
          if (mode == B200_INTRA_PLANAR) {
            const int hor = (L[y + 1] << l2w) + (x + 1) * (tr - L[y + 1]), vert = (T[x + 1] << l2h) + (y + 1) * (bl - T[x + 1]);
            v = ((hor << l2h) + (vert << l2w) + (1 << (l2w + l2h))) >> (1 + l2w + l2h);
          } else v = dc;
          v = (int16_t)v;
          if (doPDPC) {
            const int wT = 32 >> min(31, (y << 1) >> scale), wL = 32 >> min(31, (x << 1) >> scale);
            v = (int16_t)(v + ((wL * (L[y + 1] - v) + wT * (T[x + 1] - v) + 32) >> 6));
          }
          V2_STORE(x, y, v);
        }
      } else if (mode >= B200_INTRA_LM) {
        // ---- cross-component linear model, 4:2:0: luma at (lx0 + dx, ly0 + dy) read through the tile
        const bool aCu = t.flags & B200_INTRA_LM_ABOVE, lCu = t.flags & B200_INTRA_LM_LEFT, colloc = t.flags & B200_INTRA_LM_COLLOCATED;
        const int lx0 = 2 * x0, ly0 = 2 * y0;
        const bool firstRowOfCtu = ((2 * y0) & ((1 << P.ctuLog2) - 1)) == 0;
        const int nTop = aCu ? (mode == B200_INTRA_MDLM_T ? 2 * t.lmAbove : w) : 0, nLeft = lCu ? (mode == B200_INTRA_MDLM_L ? 2 * t.lmLeft : h) : 0;
        const int lox = TL.ox[0], loy = TL.oy[0], ltw = TL.tw[0], lth = TL.th[0]; const int16_t* lrec = TL.rec[0]; const int16_t* lplane = TL.plane[0]; const int lps = TL.ps[0];
        auto pixY = [&](int x, int y) -> int {
          return ((unsigned)(x - lox) < (unsigned)ltw && (unsigned)(y - loy) < (unsigned)lth) ? (int)lrec[(y - loy) * V2_LS + (x - lox)] : (int)__ldcg(lplane + (size_t)y * lps + x);
        };
#define LY(dx, dy) pixY(lx0 + (dx), ly0 + (dy))
        for (int kk = lane; kk < nTop + nLeft + w * h; kk += V2_GROUP) {
          if (kk < nTop) {
            const int i = kk, m = (i == 0 && !lCu) ? 0 : 1;
            int v;
            if (firstRowOfCtu) v = (LY(2 * i, -1) * 2 + LY(2 * i - m, -1) + LY(2 * i + 1, -1) + 2) >> 2;
            else if (colloc) v = (LY(2 * i, -3) + LY(2 * i, -2) * 4 + LY(2 * i - m, -2) + LY(2 * i + 1, -2) + LY(2 * i, -1) + 4) >> 3;
            else v = (LY(2 * i, -2) * 2 + LY(2 * i - m, -2) + LY(2 * i + 1, -2) + LY(2 * i, -1) * 2 + LY(2 * i - m, -1) + LY(2 * i + 1, -1) + 4) >> 3;
            SC.LmTop[i] = (int16_t)v;
          } else if (kk < nTop + nLeft) {
            const int j = kk - nTop;
            int v;
            if (colloc) v = (LY(-2, 2 * j - ((j == 0 && !aCu) ? 0 : 1)) + LY(-2, 2 * j) * 4 + LY(-3, 2 * j) + LY(-1, 2 * j) + LY(-2, 2 * j + 1) + 4) >> 3;
            else v = (LY(-2, 2 * j) * 2 + LY(-3, 2 * j) + LY(-1, 2 * j) + LY(-2, 2 * j + 1) * 2 + LY(-3, 2 * j + 1) + LY(-1, 2 * j + 1) + 4) >> 3;
            SC.LmLeft[j] = (int16_t)v;
          } else {
            const int q = kk - nTop - nLeft, j = q >> t.log2w, i = q & (w - 1), m = (i == 0 && !lCu) ? 0 : 1;
            int v;
            if (colloc) { const int up = (j == 0 && !aCu) ? 0 : 1; v = (LY(2 * i, 2 * j - up) + LY(2 * i, 2 * j) * 4 + LY(2 * i - m, 2 * j) + LY(2 * i + 1, 2 * j) + LY(2 * i, 2 * j + 1) + 4) >> 3; }
            else v = (LY(2 * i, 2 * j) * 2 + LY(2 * i + 1, 2 * j) + LY(2 * i - m, 2 * j) + LY(2 * i, 2 * j + 1) * 2 + LY(2 * i + 1, 2 * j + 1) + LY(2 * i - m, 2 * j + 1) + 4) >> 3;
            SC.Lm[q] = (int16_t)v;
          }
        }
#undef LY
        V2_SYNC();
        if (lane == 0) {                                            // xGetLMParameters
          const int tuWU = w >> 1, tuHU = h >> 1;
          bool aboveAvail = false, leftAvail = false; int topNum = 0, leftNum = 0;
          if (mode == B200_INTRA_MDLM_T) { aboveAvail = t.lmAbove >= tuWU; topNum = 2 * t.lmAbove; }
          else if (mode == B200_INTRA_MDLM_L) { leftAvail = t.lmLeft >= tuHU; leftNum = 2 * t.lmLeft; }
          else { aboveAvail = aCu; leftAvail = lCu; topNum = w; leftNum = h; }
          const int aboveIs4 = leftAvail ? 0 : 1, leftIs4 = aboveAvail ? 0 : 1;
          const int start0 = topNum >> (2 + aboveIs4), step0 = max(1, topNum >> (1 + aboveIs4)), start1 = leftNum >> (2 + leftIs4), step1 = max(1, leftNum >> (1 + leftIs4));
          int sl[4] = {0, 0, 0, 0}, sc[4] = {0, 0, 0, 0}, cntT = 0, cntL = 0;
          if (aboveAvail) { cntT = min(topNum, (1 + aboveIs4) << 1); for (int q = 0, pos = start0; q < cntT; q++, pos += step0) { sl[q] = SC.LmTop[pos]; sc[q] = T[1 + pos]; } }
          if (leftAvail) { cntL = min(leftNum, (1 + leftIs4) << 1); for (int q = 0, pos = start1; q < cntL; q++, pos += step1) { sl[q + cntT] = SC.LmLeft[pos]; sc[q + cntT] = L[1 + pos]; } }
          if (cntT + cntL == 2) { sl[3] = sl[0]; sc[3] = sc[0]; sl[2] = sl[1]; sc[2] = sc[1]; sl[0] = sl[1]; sc[0] = sc[1]; sl[1] = sl[3]; sc[1] = sc[3]; }
          int mn0 = 0, mn1 = 2, mx0 = 1, mx1 = 3, tt;
          if (sl[mn0] > sl[mn1]) { tt = mn0; mn0 = mn1; mn1 = tt; }
          if (sl[mx0] > sl[mx1]) { tt = mx0; mx0 = mx1; mx1 = tt; }
          if (sl[mn0] > sl[mx1]) { tt = mn0; mn0 = mx0; mx0 = tt; tt = mn1; mn1 = mx1; mx1 = tt; }
          if (sl[mn1] > sl[mx0]) { tt = mn1; mn1 = mx0; mx0 = tt; }
          const int minL = (sl[mn0] + sl[mn1] + 1) >> 1, minC = (sc[mn0] + sc[mn1] + 1) >> 1, maxL = (sl[mx0] + sl[mx1] + 1) >> 1, maxC = (sc[mx0] + sc[mx1] + 1) >> 1;
          int a = 0, b = 1 << (P.bitDepth - 1), shift = 0;
          if (leftAvail || aboveAvail) {
            const int diff = maxL - minL;
            if (diff > 0) {
              const int diffC = maxC - minC;
              int x = 31 - __clz(diff);
              const int normDiff = (diff << 4 >> x) & 15;
              const int v = (int)((0x0765544332211110ull >> (4 * (15 - normDiff))) & 15) | 8;
              x += normDiff != 0;
              const int y = diffC == 0 ? 0 : (31 - __clz(abs(diffC))) + 1, add = 1 << y >> 1;
              a = (diffC * v + add) >> y; shift = 3 + x - y;
              if (shift < 1) { shift = 1; a = a == 0 ? 0 : a < 0 ? -15 : 15; }
              b = minC - ((a * minL) >> shift);
            } else { a = 0; b = minC; shift = 0; }
          }
          SC.LmPar[0] = a; SC.LmPar[1] = b; SC.LmPar[2] = shift;
        }
        V2_SYNC();
        const int a = SC.LmPar[0], b = SC.LmPar[1], shift = SC.LmPar[2];
        for (int kk = lane; kk < w * h; kk += V2_GROUP) { const int y = kk >> t.log2w, x = kk & (w - 1); V2_STORE(x, y, clip3(0, pmax, ((a * SC.Lm[kk]) >> shift) + b)); }
      } else if (mode == B200_INTRA_MIP) {
        const int sizeId = (w == 4 && h == 4) ? 0 : (w == 4 || h == 4 || (w == 8 && h == 8)) ? 1 : 2;
        const int bdry = sizeId == 0 ? 2 : 4, red = sizeId < 2 ? 4 : 8, upH = w / red, upV = h / red, inSize = 2 * bdry;
        const int modeIdx = t.mip & 0x7f; const bool transpose = t.mip >> 7;
        int16_t* in = SC.S; int16_t* sM = SC.M;
        if (lane < inSize) {
          const bool fromLeft = (lane >= bdry) != transpose; const int d = lane % bdry, len = fromLeft ? h : w;
          const int16_t* full = (fromLeft ? L : T) + 1;
          int v;
          if (bdry < len) { const int f = len / bdry; int sum = 0; for (int j = 0; j < f; j++) sum += full[d * f + j]; v = (sum + (f >> 1)) >> (31 - __clz(f)); }
          else v = full[d];
          in[lane] = (int16_t)v;
        }
        V2_SYNC();
        const int inputOffset = in[0];
        V2_SYNC();
        if (lane < inSize) in[lane] = (int16_t)(lane == 0 ? (sizeId < 2 ? (1 << (P.bitDepth - 1)) - inputOffset : 0) : in[lane] - inputOffset);
        V2_SYNC();
        for (int o = lane; o < red * red; o += V2_GROUP) {
          const int redSize = sizeId == 2, stride = inSize - redSize;
          const uint8_t* wgt = (sizeId == 0 ? kMip4x4 + modeIdx * 64 : sizeId == 1 ? kMip8x8 + modeIdx * 128 : kMip16x16 + modeIdx * 448) + o * stride;
          int sum = 0, acc = 0;
          for (int i = 0; i < inSize; i++) sum += in[i];
          for (int i = redSize; i < inSize; i++) acc += in[i] * wgt[i - redSize];
          const int v = clip3(0, pmax, ((acc + 32 - 32 * sum) >> 6) + inputOffset);
          sM[transpose ? (o % red) * red + o / red : o] = (int16_t)v;
        }
        V2_SYNC();
        const int l2H = 31 - __clz(upH), l2V = 31 - __clz(upV);
        for (int kk = lane; kk < w * h; kk += V2_GROUP) {
          const int y = kk >> t.log2w, x = kk & (w - 1), kr = y / upV, i = y % upV;
          int hv[2];
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const int k2 = kr - 1 + q;
            if (k2 < 0) hv[q] = T[x + 1];
            else if (upH == 1) hv[q] = sM[k2 * red + x];
            else {
              const int j = x / upH, ii = x % upH, before = j == 0 ? L[(k2 + 1) * upV] : sM[k2 * red + j - 1], behind = sM[k2 * red + j];
              hv[q] = (int16_t)(before * upH + (upH >> 1) + (ii + 1) * (behind - before)) >> l2H;
            }
          }
          V2_STORE(x, y, upV == 1 ? hv[1] : (int16_t)(hv[0] * upV + (upV >> 1) + (i + 1) * (hv[1] - hv[0])) >> l2V);
        }
      } else if (mode >= B200_INTRA_BDPCM_HOR) {
        for (int kk = lane; kk < w * h; kk += V2_GROUP) { const int y = kk >> t.log2w, x = kk & (w - 1); V2_STORE(x, y, mode == B200_INTRA_BDPCM_HOR ? L[y + 1] : T[x + 1]); }
      } else {
        // ---- angular (xPredIntraAng)
        const int predMode = wide_angle(bw, bh, mode);                 // ISP: the CU decides (:610)
        const bool ver = predMode >= 34;
        const int angMode = ver ? predMode - 50 : -(predMode - 18), absMode = abs(angMode);
        const int invAngle = cInvAng[absMode], absAng = cAng[absMode], angle = angMode < 0 ? -absAng : absAng;
        const int16_t *mainSrc = ver ? T : L, *sideSrc = ver ? L : T;
        const int mw = ver ? w : h, mh = ver ? h : w;
        int16_t *M = SC.M + IT_ORG, *S = SC.S + IT_ORG;
        if (angle < 0) {
          for (int kk = lane - mh; kk <= mw + 1 + mrl; kk += V2_GROUP) M[kk] = kk >= 0 ? mainSrc[kk] : sideSrc[min((-kk * invAngle + 256) >> 9, mh)];
          for (int kk = lane; kk <= mh + 1 + mrl; kk += V2_GROUP) S[kk] = sideSrc[kk];
        } else {
          const int l2r = (31 - __clz(mw)) - (31 - __clz(mh)), sft = max(0, l2r), maxIndex = (mrl << sft) + 2, refLength = ver ? topLen : sideLen;
          for (int kk = lane; kk <= refLength + mrl + maxIndex; kk += V2_GROUP) M[kk] = mainSrc[min(kk, refLength + mrl)];
          for (int kk = lane; kk <= (ver ? sideLen : topLen) + mrl; kk += V2_GROUP) S[kk] = sideSrc[kk];
        }
        V2_SYNC();
        const int16_t *Mp = M + mrl, *Sp = S + mrl;
        const int l2mw = 31 - __clz(mw), l2mh = 31 - __clz(mh);
        const int topLeft = T[0];
        const int scale0 = (l2mw - 2 + l2mh - 2 + 2) >> 2;
        const int lev = min(3 << scale0, mw);
        const bool frac = (absAng & 31) != 0;
        const int diff = min(abs(predMode - 18), abs(predMode - 50));
        const bool cubic = isp || !(diff > cIntraFilterThr[(l2mw + l2mh) >> 1]) || mrl > 0;
        int angularScale = -1;
        if (angle > 0 && doPDPC) angularScale = min(2, l2mh - ((31 - __clz(3 * invAngle - 2)) - 8));
#pragma unroll 2
        for (int kk = lane; kk < mw * mh; kk += V2_GROUP) {
          const int yy = kk >> l2mw, xx = kk & (mw - 1);
          int v;
          if (angle == 0) {
            if (doPDPC && xx < lev) { const int wL = 32 >> min(31, (xx << 1) >> scale0); v = clip3(0, pmax, (wL * (Sp[yy + 1] - topLeft) + Mp[xx + 1] * 64 + 32) >> 6); }
            else v = Mp[xx + 1];
          } else {
            const int deltaPos = angle * (1 + mrl + yy), dI = deltaPos >> 5, dF = deltaPos & 31;
            if (!frac) v = Mp[dI + 1 + xx];
            else if (c) v = (int16_t)(((32 - dF) * Mp[dI + 1 + xx] + dF * Mp[dI + 2 + xx] + 16) >> 5);
            else {
              const int16_t* p = Mp + dI + xx;
              int f0, f1, f2, f3;
              if (cubic) { f0 = kIfChroma[dF * 4]; f1 = kIfChroma[dF * 4 + 1]; f2 = kIfChroma[dF * 4 + 2]; f3 = kIfChroma[dF * 4 + 3]; }
              else { f0 = 16 - (dF >> 1); f1 = 32 - (dF >> 1); f2 = 16 + (dF >> 1); f3 = dF >> 1; }
              v = (int16_t)((f0 * p[0] + f1 * p[1] + f2 * p[2] + f3 * p[3] + 32) >> 6);
              if (cubic) v = clip3(0, pmax, v);
            }
            if (angularScale >= 0 && xx < min(3 << angularScale, mw)) {
              const int invAngleSum = 256 + (xx + 1) * invAngle, wL = 32 >> (2 * xx >> angularScale), left = Sp[yy + (invAngleSum >> 9) + 1];
              v = (int16_t)(v + ((wL * (left - v) + 32) >> 6));
            }
          }
          if (ver) V2_STORE(xx, yy, v); else V2_STORE(yy, xx, v);
        }
      }
#undef V2_STORE
      // ---- publish: the CTU's own later blocks see the tile; other CTUs read the plane, and only the samples of the CTU's last column and last row
      // (left, above and above-right references of the CTUs right of / below it): blocks that touch neither skip the device-scope fence and the done word
      __threadfence_block();
      V2_SYNC();
      K6P(pb + 3, tp); tp = clock64();                          // [3] prediction + stores
      if (lane == 0) sflag[k] = 1;
      if (x0 + w == tox + ttw || y0 + h == toy + tth) {
        __threadfence();
        V2_SYNC();
        if (lane == 0) atomicExch(P.done + me, 1);
      }
      K6P(pb + 4, tp); K6C(pb + 9); K6P(pb + 10 + min(4, max(0, ((int)t.log2w + (int)t.log2h - 4) >> 1)), tb0);   // [4] device fence + done word, [9] blocks, [10..14] whole block by size class
    }
#undef V2_SYNC
  }
}

// record checks of the picture path (the kernel-level wrapper checks on the host): everything K6 uses as an address
__global__ void __launch_bounds__(256) intra_validate_kernel(const b200_intra_tu* __restrict__ tus, int n, int W, int H, int chroma, int* meta)
{
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const b200_intra_tu t = tus[i];
  if (t.flags & B200_INTRA_ISP) { b200_intra_tu prev; if (i) prev = tus[i - 1]; if (!intra_isp_record_ok(t, i ? &prev : nullptr, W, H)) atomicOr(&meta[LM_ERR], 8); return; }
  const int w = 1 << t.log2w, h = 1 << t.log2h, pw = t.comp ? W >> 1 : W, ph = t.comp ? H >> 1 : H, unit = t.comp ? 2 : 4, m = t.multiRefIdx;
  bool ok = t.comp < (chroma ? 3 : 1) && t.log2w >= 2 && t.log2w <= 6 && t.log2h >= 1 && t.log2h <= 6 && t.x + w <= pw && t.y + h <= ph && !(t.x % unit) && !(t.y % unit);
  ok = ok && t.mode <= B200_INTRA_MDLM_T && m <= 2 && (!m || !t.comp);
  if (t.ciip) ok = ok && t.ciip <= 3 && t.mode == B200_INTRA_PLANAR;
  if (t.mode >= B200_INTRA_LM) ok = ok && t.comp && t.log2w <= 5 && t.log2h <= 5 && t.lmAbove <= w && t.lmLeft <= h && (!(t.flags & B200_INTRA_LM_ABOVE) || t.y >= 2) && (!(t.flags & B200_INTRA_LM_LEFT) || t.x >= 2)
                                  && t.x + max(w, 2 * t.lmAbove) <= pw && t.y + max(h, 2 * t.lmLeft) <= ph;
  if (t.mode == B200_INTRA_MIP) ok = ok && !t.comp && !m && (t.mip & 0x7f) < ((w == 4 && h == 4) ? 16 : (w == 4 || h == 4 || (w == 8 && h == 8)) ? 8 : 6);
  ok = ok && t.numAbove <= 2 * w / unit && t.numLeft <= 2 * h / unit && (!t.numAbove || t.y > m) && (!t.numLeft || t.x > m)
          && (!(t.flags & B200_INTRA_AVAIL_TL) || (t.x > m && t.y > m)) && t.x + (int)t.numAbove * unit <= pw && t.y + (int)t.numLeft * unit <= ph;
  if (!ok) atomicOr(&meta[LM_ERR], 8);
}

#ifdef B200_K6_PROF
extern "C" __attribute__((visibility("default"))) void b200_k6_prof_dump(int reset)
{
  unsigned long long h[48]; cudaDeviceSynchronize(); cudaMemcpyFromSymbol(h, gK6Prof, sizeof(h));
  const double nc = (double)(h[8] ? h[8] : 1);
  fprintf(stderr, "K6 prof: %.0f group-CTUs, set-up %.0f cyc each\n", nc, h[0] / nc);
  for (int c = 0; c < 2; c++) {
    const int b = c * 16; const double nb = (double)(h[b + 9] ? h[b + 9] : 1);
    fprintf(stderr, "  %s: %.0f blocks | wait %.0f, refs %.0f, predict %.0f, publish %.0f cyc | whole block by size class (<=32, <=128, <=512, <=2048, more samples): %.0f %.0f %.0f %.0f %.0f cyc-sums/blocks\n", c ? "chroma" : "luma", nb,
            h[b + 1] / nb, h[b + 2] / nb, h[b + 3] / nb, h[b + 4] / nb, h[b + 10] / nb, h[b + 11] / nb, h[b + 12] / nb, h[b + 13] / nb, h[b + 14] / nb);
  }
  if (reset) { memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(gK6Prof, h, sizeof(h)); }
}
#endif
int launch_intra_validate(const b200_intra_tu* tus, size_t numTus, const b200_geom& g, int* meta, cudaStream_t s)
{
  if (!numTus) return 0;
  intra_validate_kernel<<<(unsigned)((numTus + 255) / 256), 256, 0, s>>>(tus, (int)numTus, g.width, g.height, g.chromaFormat != 0, meta);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int launch_intra(const IntraLaunch& L, cudaStream_t s)
{
  if (!L.numTus) return 0;
  IntraParams P;
  for (int c = 0; c < 3; c++) { P.planes[c] = L.planes.p[c]; P.resi[c] = L.resi[c]; P.stride[c] = L.planes.stride[c]; P.owner[c] = L.owner[c]; P.ownerStride[c] = L.ownerStride[c]; }
  if (!L.geom.chromaFormat) { P.planes[1] = P.planes[2] = nullptr; }
  P.W = L.geom.width; P.H = L.geom.height; P.bitDepth = L.geom.bitDepth; P.tus = L.tus; P.numTus = (int)L.numTus;
  P.done = L.sync; P.ticket = L.sync + L.numTus; P.err = L.sync + L.numTus + 1; P.compSel = L.compSel;
  const bool cont = L.compSel == 2;                            // second launch on the list: done words, owner maps and the processing order are kept
  if (cont) B200_CUDA(cudaMemsetAsync(P.ticket, 0, sizeof(int), s));
  else {
    B200_CUDA(cudaMemsetAsync(L.sync, 0, (L.numTus + 2) * sizeof(int), s));
    for (int c = 0; c < (L.geom.chromaFormat ? 3 : 1); c++) B200_CUDA(cudaMemsetAsync(L.owner[c], 0xff, L.ownerBytes[c], s));
  }
  P.perm = nullptr; P.ctuCnt = P.ctuFirst = P.ctuBase = nullptr;
  P.ctuLog2 = L.geom.ctuSize == 128 ? 7 : L.geom.ctuSize == 64 ? 6 : 5; P.ctusW = (L.geom.width + L.geom.ctuSize - 1) / L.geom.ctuSize; P.ctusH = (L.geom.height + L.geom.ctuSize - 1) / L.geom.ctuSize;
  static const char* variant = getenv("B200_INTRA_KERNEL");      // measurement switch: "v1" = one CTA per block through global memory (round 1)
  // v2 (CTU-resident) shortens the dependency chains of dense lists (I pictures); the scattered intra CUs of a B picture have almost no chains and
  // finish sooner with v1's one-CTA-per-block throughput (measured at 4K: 15 % intra CUs 0.08 ms vs 0.24 ms; I picture 12.5 ms vs 8 ms)
  const bool dense = L.numTus >= 48 * std::max<size_t>(1, (size_t)L.geom.width * L.geom.height >> 14);                   // >= 48 blocks per 128x128 luma area
  const bool v1 = !L.order || (variant ? !strcmp(variant, "v1") : !dense) || ((P.stride[0] | P.stride[1] | P.stride[2]) & 1);   // the tile loads move 32-bit words
  const unsigned grid = (unsigned)((L.numTus + 255) / 256);
  if (!cont) intra_owner_kernel<<<(unsigned)L.numTus, 64, 0, s>>>(P);
  if (!v1) {
    // v2: per-CTU runs of the list (decoding order keeps a CTU's blocks together), CTUs in wave-front order, one CTA per CTU at a time
    const size_t nCtu = (size_t)P.ctusW * P.ctusH;
    int* base = L.order; P.ctuCnt = base; P.ctuFirst = base + nCtu; int* ctuOrder = base + 2 * nCtu; int* counters = base + 3 * nCtu;
    if (cont) B200_CUDA(cudaMemsetAsync(counters + 1, 0, sizeof(int), s));
    else {
      B200_CUDA(cudaMemsetAsync(P.ctuCnt, 0, nCtu * sizeof(int), s));
      B200_CUDA(cudaMemsetAsync(P.ctuFirst, 0x7f, nCtu * sizeof(int), s));
      intra_ctu_count_kernel<<<grid, 256, 0, s>>>(P);
      intra_ctu_check_kernel<<<grid, 256, 0, s>>>(P);
      intra_ctu_order_kernel<<<1, 1024, nCtu * sizeof(int), s>>>(P, ctuOrder, counters);
    }
    const int ctas = (int)std::min<size_t>(nCtu, (size_t)num_sms());
    static const int shape = getenv("B200_INTRA_GROUP") ? atoi(getenv("B200_INTRA_GROUP")) : 0;      // measurement switch: threads per block group
#define V2_GO(G, N) do { static bool attr = false; if (!attr) { B200_CUDA(cudaFuncSetAttribute(intra_ctu_kernel<G, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2_smem(N))); attr = true; } \
                         intra_ctu_kernel<G, N><<<ctas, G * N, v2_smem(N), s>>>(P, ctuOrder, counters); } while (0)
    // measured on a 4K I picture (66810 blocks): 128x4 7.45 ms, 128x6 6.97 ms, 128x8 7.70 ms, 64x6 7.62 ms, 64x12 7.61 ms, 256x3 9.22 ms, 32x8 17.9 ms
    if (shape == 1284) V2_GO(128, 4); else if (shape == 32) V2_GO(32, 8); else V2_GO(128, 6);
#undef V2_GO
    B200_CUDA(cudaGetLastError());
    return 0;
  }
  static const bool listOrder = getenv("B200_INTRA_ORDER") && !strcmp(getenv("B200_INTRA_ORDER"), "decode");   // measurement switch: tickets in list order
  if (L.order && !listOrder && L.numTus >= 100 * std::max<size_t>(1, (size_t)L.geom.width * L.geom.height >> 14)) {   // >= 100 blocks per 128x128 luma area
    const size_t nCtu = (size_t)P.ctusW * P.ctusH;
    int* perm = L.order; P.ctuCnt = perm + L.numTus; P.ctuFirst = P.ctuCnt + nCtu; P.ctuBase = P.ctuFirst + nCtu;
    if (!cont) {
      B200_CUDA(cudaMemsetAsync(P.ctuCnt, 0, nCtu * sizeof(int), s));
      B200_CUDA(cudaMemsetAsync(P.ctuFirst, 0x7f, nCtu * sizeof(int), s));
      intra_ctu_count_kernel<<<grid, 256, 0, s>>>(P);
      intra_ctu_base_kernel<<<1, 1, 0, s>>>(P);
      intra_perm_kernel<<<grid, 256, 0, s>>>(P, perm);
    }
    P.perm = perm;
  }
  const int ctas = (int)std::min<size_t>(L.numTus, (size_t)num_sms() * 12);
  intra_kernel<<<ctas, IT_THREADS, 0, s>>>(P);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200
