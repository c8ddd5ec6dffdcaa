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
#define VVC_TABLE_QUAL static __device__ const
#include "vvc_tables.h"
#include "common.cuh"

namespace b200 {

constexpr int IFO = 8192;   // IF_INTERNAL_OFFS

struct McParams {
  int16_t* dst[3]; int dstStride[3];
  const int16_t* refs[B200_MAX_SLOTS * 3];   // device plane pointers, by value (no table in memory -> no sync when the DPB mapping changes)
  int refStride[3];
  int W, H, bitDepth, ctuSize, chroma;
  const b200_pu* pus; const uint32_t* tiles; int numTiles;
  int32_t* dmvrMv;
};

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

__device__ __forceinline__ int avg_bi(int p0, int p1, int w1, int hr, int pmax)
{
  int v;
  if (w1 == 4) v = (p0 + p1 + (1 << hr) + 2 * IFO) >> (hr + 1);
  else         v = (p0 * (8 - w1) + p1 * w1 + (1 << (hr + 2)) + (IFO << 3)) >> (hr + 3);
  return clip3(0, pmax, v);
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
constexpr int WS = 24;   // luma window row stride (23 used)

__global__ void __launch_bounds__(256) mc_tile_kernel(const McParams P)
{
  __shared__ int16_t sW[2][23 * WS];        // luma windows / (DMVR) bilinear buffers 20x20
  __shared__ int16_t sHf[2][23 * 16];       // after horizontal filter
  __shared__ int16_t sP[2][18 * 18];        // 14-bit predictions with 1-sample ring
  __shared__ int16_t sG[2][2][18 * 18];     // gradX / gradY
  __shared__ int     sVxy[16][2];
  __shared__ unsigned sSad[25];
  __shared__ int     sDec[4];               // dmvX, dmvY, bio, -
  __shared__ int16_t sCW[4][11 * 12], sCH[4][11 * 8], sCP[4][64];

  const int tid = threadIdx.x;
  const uint32_t tile = P.tiles[blockIdx.x];
  const b200_pu pu = P.pus[tile >> 6];
  const int tx0 = (tile & 7) * 16, ty0 = ((tile >> 3) & 7) * 16;
  const int tw = min(16, pu.w - tx0), th = min(16, pu.h - ty0);
  const int bx = pu.x + tx0, by = pu.y + ty0;              // tile position (luma)
  const int bd = P.bitDepth, pmax = (1 << bd) - 1, hr = max(2, 14 - bd), sh1 = 6 - hr;
  const bool bi = pu.refSlot[0] >= 0 && pu.refSlot[1] >= 0;
  const bool altHpel = pu.flags & B200_PU_ALTHPEL;
  const bool dmvr = pu.flags & B200_PU_DMVR;
  const bool is4x4 = pu.w == 4 && pu.h == 4;   // InterpolationFilter::filterHor/Ver pick the 6-tap table for 4x4 blocks (:1062,:1155)
  bool bio = pu.flags & B200_PU_BDOF;
  const int nList = bi ? 2 : 1, l0 = pu.refSlot[0] >= 0 ? 0 : 1;

  RefPl R[2][3];
#pragma unroll
  for (int l = 0; l < 2; l++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int slot = pu.refSlot[l] < 0 ? 0 : pu.refSlot[l];
      R[l][c].p = P.refs[slot * 3 + c]; R[l][c].w = c ? P.W >> 1 : P.W; R[l][c].h = c ? P.H >> 1 : P.H; R[l][c].stride = P.refStride[c];
    }

  int mv[2][2];                                             // MV used for the final MC (clipped)
  Win win[2][2];                                            // [list][luma|chroma] access windows
  int org[2][2][2];                                         // [list][luma|chroma][x|y]: integer reference position of output sample (0,0)
#pragma unroll
  for (int l = 0; l < 2; l++) {
    mv[l][0] = pu.mv[l][0]; mv[l][1] = pu.mv[l][1];
    clip_mv(mv[l][0], mv[l][1], pu.x, pu.y, P);            // relative to the CU (xPredInterUni :651; xinitMC :1811)
#pragma unroll
    for (int k = 0; k < 2; k++) { win[l][k].x0 = win[l][k].y0 = -(1 << 20); win[l][k].x1 = win[l][k].y1 = 1 << 20; }
  }

  // ================================================================ DMVR search (xProcessDMVR :1847)
  if (dmvr) {
    // bilinear 10-bit predictions (tw+4)x(th+4) around the clipped merge MV - 2 (xinitMC :1804)
    for (int i = tid; i < 2 * (tw + 4) * (th + 4); i += 256) {
      const int l = i / ((tw + 4) * (th + 4)), j = i - l * (tw + 4) * (th + 4);
      const int y = j / (tw + 4), x = j - y * (tw + 4);
      // the search buffer of the whole CU starts at CU + mv - 2; this tile's part starts tx0,ty0 further
      const int mx = mv[l][0] - 32, my = mv[l][1] - 32;
      const int xF = mx & 15, yF = my & 15, X = pu.x + tx0 + (mx >> 4) + x, Y = pu.y + ty0 + (my >> 4) + y;
      const int8_t* fh = kIfBilin4 + xF * 2; const int8_t* fv = kIfBilin4 + yF * 2;
      const int s1 = 4 - (10 - bd), o1 = 1 << (s1 - 1);
      const RefPl& r = R[l][0];
      int v;
      if (xF == 0 && yF == 0) v = ldc(r, X, Y) << (10 - bd);
      else if (yF == 0) v = (fh[0] * ldc(r, X, Y) + fh[1] * ldc(r, X + 1, Y) + o1) >> s1;
      else if (xF == 0) v = (fv[0] * ldc(r, X, Y) + fv[1] * ldc(r, X, Y + 1) + o1) >> s1;
      else {
        const int a = (int16_t)((fh[0] * ldc(r, X, Y) + fh[1] * ldc(r, X + 1, Y) + o1) >> s1);
        const int b = (int16_t)((fh[0] * ldc(r, X, Y + 1) + fh[1] * ldc(r, X + 1, Y + 1) + o1) >> s1);
        v = (fv[0] * a + fv[1] * b + 8) >> 4;
      }
      sW[l][y * WS + x] = (int16_t)v;
    }
    if (tid < 25) sSad[tid] = 0;
    __syncthreads();
    if (tid < 200) {                                       // SAD over every second row (RdCost.cpp:113-135), 25 positions x 8 rows
      const int p = tid >> 3, y = (tid & 7) * 2;
      if (y < th) {
        const int u = p % 5 - 2, v = p / 5 - 2;
        const int16_t* a = &sW[0][(2 + v + y) * WS + 2 + u]; const int16_t* b = &sW[1][(2 - v + y) * WS + 2 - u];
        unsigned s = 0;
        for (int x = 0; x < tw; x++) s += abs(a[x] - b[x]);
        atomicAdd(&sSad[p], s);
      }
    }
    __syncthreads();
    if (tid == 0) {
      unsigned sads[25];
      for (int i = 0; i < 25; i++) sads[i] = sSad[i];
      unsigned minCost = sads[12]; minCost -= minCost >> 2;  // (:1924-1925)
      int dx = 0, dy = 0;
      if (minCost >= (unsigned)(tw * th)) {
        sads[12] = minCost;
        int bu = 0, bv = 0;
        for (int i = 0; i < 25; i++) if (sads[i] < minCost) { minCost = sads[i]; bu = i % 5 - 2; bv = i / 5 - 2; }   // xBIPMVRefine, raster order, strict <
        dx = bu * 16; dy = bv * 16;
        if (abs(dx) != 32 && abs(dy) != 32) {              // xDMVRSubPixelErrorSurface / xSubPelErrorSrfc
          const unsigned* c = &sads[(bv + 2) * 5 + bu + 2];
          const unsigned long long s0 = c[0], sl = c[-1], st = c[-5], sr = c[1], sb = c[5];
          {
            const long long num = (long long)(sl - sr) * 16, den = (long long)(sl + sr - (s0 << 1));
            if (den != 0) dx += (sl != s0 && sr != s0) ? div_for_maxq7(num, den) : (sl == s0 ? -8 : 8);
          }
          {
            const long long num = (long long)(st - sb) * 16, den = (long long)(st + sb - (s0 << 1));
            if (den != 0) dy += (st != s0 && sb != s0) ? div_for_maxq7(num, den) : (st == s0 ? -8 : 8);
          }
        }
      }
      sDec[0] = dx; sDec[1] = dy;
      sDec[2] = (minCost < (unsigned)(2 * tw * th)) ? 0 : 1;  // bioAppliedSubblk (:1984)
      if (P.dmvrMv) {
        const int num = (ty0 >> 4) * max(1, pu.w >> 4) + (tx0 >> 4);
        P.dmvrMv[(pu.dmvrOff + num) * 2] = dx; P.dmvrMv[(pu.dmvrOff + num) * 2 + 1] = dy;
      }
    }
    __syncthreads();
    bio = bio && sDec[2];
    const int dmx = sDec[0], dmy = sDec[1];
#pragma unroll
    for (int l = 0; l < 2; l++) {
      const int mrgx = pu.mv[l][0], mrgy = pu.mv[l][1];
      const int rx = clip3(-(1 << 17), (1 << 17) - 1, l ? mrgx - dmx : mrgx + dmx), ry = clip3(-(1 << 17), (1 << 17) - 1, l ? mrgy - dmy : mrgy + dmy);
      int cx = rx, cy = ry;
      clip_mv(cx, cy, bx, by, P);                          // cMvClipped, relative to the sub-block (:1749)
      mv[l][0] = cx; mv[l][1] = cy;
#pragma unroll
      for (int k = 0; k < 2; k++) {                        // k = 0 luma, 1 chroma (xFinalPaddedMCForDMVR :1757-1778, xPrefetchPad :1525)
        const int sh = 4 + k, taps = k ? 4 : 8, cs = k;
        const int dIx = (rx >> sh) - (mrgx >> sh), dIy = (ry >> sh) - (mrgy >> sh);
        if (dIx || dIy) {
          int pmx = mrgx - ((taps / 2 - 1) << sh), pmy = mrgy - ((taps / 2 - 1) << sh);
          clip_mv(pmx, pmy, bx, by, P);
          Win w; w.x0 = (bx >> cs) + (pmx >> sh); w.y0 = (by >> cs) + (pmy >> sh);
          w.x1 = w.x0 + (tw >> cs) + taps - 2; w.y1 = w.y0 + (th >> cs) + taps - 2;
          win[l][k] = w;
          org[l][k][0] = w.x0 + (taps / 2 - 1) + dIx; org[l][k][1] = w.y0 + (taps / 2 - 1) + dIy;
        } else {
          org[l][k][0] = (bx >> cs) + (cx >> sh); org[l][k][1] = (by >> cs) + (cy >> sh);
        }
      }
    }
  } else {
#pragma unroll
    for (int l = 0; l < 2; l++) {
      org[l][0][0] = bx + (mv[l][0] >> 4); org[l][0][1] = by + (mv[l][1] >> 4);
      org[l][1][0] = (bx >> 1) + (mv[l][0] >> 5); org[l][1][1] = (by >> 1) + (mv[l][1] >> 5);
    }
  }
  if (!bi) bio = false;

  // ================================================================ luma: stage windows, H filter, V filter
  __syncthreads();
  for (int i = tid; i < nList * (tw + 7) * (th + 7); i += 256) {
    const int li = i / ((tw + 7) * (th + 7)), j = i - li * (tw + 7) * (th + 7);
    const int l = bi ? li : l0;
    const int y = j / (tw + 7), x = j - y * (tw + 7);
    sW[l][y * WS + x] = (int16_t)ldw(R[l][0], win[l][0], org[l][0][0] + x - 3, org[l][0][1] + y - 3);
  }
  __syncthreads();
  for (int i = tid; i < nList * (th + 7) * tw; i += 256) {
    const int li = i / ((th + 7) * tw), j = i - li * (th + 7) * tw;
    const int l = bi ? li : l0;
    const int y = j / tw, x = j - y * tw;
    const int xF = mv[l][0] & 15;
    int s;
    if (xF == 0) s = 64 * sW[l][y * WS + x + 3];
    else {
      const int8_t* f = luma_taps(xF, is4x4, altHpel);
      s = 0;
#pragma unroll
      for (int t = 0; t < 8; t++) s += f[t] * sW[l][y * WS + x + t];
    }
    sHf[l][y * 16 + x] = (int16_t)((s - (IFO << sh1)) >> sh1);
  }
  __syncthreads();
  {
    const int y = tid >> 4, x = tid & 15;
    if (x < tw && y < th) {
      int pred[2] = { 0, 0 };
      for (int li = 0; li < nList; li++) {
        const int l = bi ? li : l0;
        const int yF = mv[l][1] & 15;
        int s;
        if (yF == 0) s = 64 * sHf[l][(y + 3) * 16 + x];
        else {
          const int8_t* f = luma_taps(yF, is4x4, altHpel);
          s = 0;
#pragma unroll
          for (int t = 0; t < 8; t++) s += f[t] * sHf[l][(y + t) * 16 + x];
        }
        pred[li] = s;
      }
      int16_t* d = P.dst[0] + (size_t)(by + y) * P.dstStride[0] + bx + x;
      if (!bi) *d = (int16_t)clip3(0, pmax, (pred[0] + (1 << (5 + hr)) + (IFO << 6)) >> (6 + hr));
      else if (!bio) *d = (int16_t)avg_bi((int16_t)(pred[0] >> 6), (int16_t)(pred[1] >> 6), pu.bcwW1, hr, pmax);
      else { sP[0][(y + 1) * 18 + x + 1] = (int16_t)(pred[0] >> 6); sP[1][(y + 1) * 18 + x + 1] = (int16_t)(pred[1] >> 6); }
    }
  }

  // ================================================================ BDOF (applyBiOptFlow :1290)
  if (bio) {
    // ring of integer reference samples around the block (xPredInterBlk :847-885); the windows already hold them
    for (int i = tid; i < 2 * 2 * (tw + th + 2); i += 256) {
      const int l = i / (2 * (tw + th + 2)), j = i - l * 2 * (tw + th + 2);
      int x, y;
      if (j < tw + 2) { x = j; y = 0; } else if (j < 2 * (tw + 2)) { x = j - (tw + 2); y = th + 1; }
      else if (j < 2 * (tw + 2) + th) { x = 0; y = j - 2 * (tw + 2) + 1; } else { x = tw + 1; y = j - 2 * (tw + 2) - th + 1; }
      const int xo = (mv[l][0] & 15) < 8 ? 1 : 0, yo = (mv[l][1] & 15) < 8 ? 1 : 0;
      // P(x,y) <- window sample at output position (x-1-xo+..): window origin = output(-3,-3)
      const int v = sW[l][(y - yo + 3) * WS + (x - xo + 3)];
      sP[l][y * 18 + x] = (int16_t)((int16_t)(v << hr) - IFO);
    }
    __syncthreads();
    // gradients on the interior (gradFilterCore<true> :212), then replicate gradients AND predictions into the ring (:236-266)
    {
      const int y = tid >> 4, x = tid & 15;
      if (x < tw && y < th) {
#pragma unroll
        for (int l = 0; l < 2; l++) {
          const int16_t* p = &sP[l][(y + 1) * 18 + x + 1];
          sG[l][0][(y + 1) * 18 + x + 1] = (int16_t)((p[1] >> 6) - (p[-1] >> 6));
          sG[l][1][(y + 1) * 18 + x + 1] = (int16_t)((p[18] >> 6) - (p[-18] >> 6));
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < 6 * 2 * th; i += 256) {          // left / right columns of 6 arrays
      const int a = i / (2 * th), j = i - a * 2 * th, y = (j >> 1) + 1, right = j & 1;
      int16_t* A = a < 2 ? sP[a] : sG[(a - 2) >> 1][(a - 2) & 1];
      if (right) A[y * 18 + tw + 1] = A[y * 18 + tw]; else A[y * 18] = A[y * 18 + 1];
    }
    __syncthreads();
    for (int i = tid; i < 6 * 2 * (tw + 2); i += 256) {    // top / bottom rows (incl. corners)
      const int a = i / (2 * (tw + 2)), j = i - a * 2 * (tw + 2), x = j >> 1, bottom = j & 1;
      int16_t* A = a < 2 ? sP[a] : sG[(a - 2) >> 1][(a - 2) & 1];
      if (bottom) A[(th + 1) * 18 + x] = A[th * 18 + x]; else A[x] = A[18 + x];
    }
    __syncthreads();
    // per 4x4 block: sums over the 6x6 window (calcBIOSums :134); 16 threads per block
    {
      const int blk = tid >> 4, k = tid & 15;
      const int bxx = (blk & 3) * 4, byy = (blk >> 2) * 4;
      int sAX = 0, sAY = 0, sDX = 0, sDY = 0, sS = 0;
      if (bxx < tw && byy < th) {
        for (int j = k; j < 36; j += 16) {
          const int yy = j / 6, xx = j - yy * 6, i = (byy + yy) * 18 + bxx + xx;
          const int gX = (sG[0][0][i] + sG[1][0][i]) >> 1, gY = (sG[0][1][i] + sG[1][1][i]) >> 1;
          const int dI = (sP[1][i] >> 4) - (sP[0][i] >> 4);
          sAX += abs(gX); sAY += abs(gY);
          sDX += gX < 0 ? -dI : (gX == 0 ? 0 : dI);
          sDY += gY < 0 ? -dI : (gY == 0 ? 0 : dI);
          sS  += gY < 0 ? -gX : (gY == 0 ? 0 : gX);
        }
      }
#pragma unroll
      for (int m = 1; m < 16; m <<= 1) {
        sAX += __shfl_xor_sync(0xffffffffu, sAX, m); sAY += __shfl_xor_sync(0xffffffffu, sAY, m);
        sDX += __shfl_xor_sync(0xffffffffu, sDX, m); sDY += __shfl_xor_sync(0xffffffffu, sDY, m);
        sS  += __shfl_xor_sync(0xffffffffu, sS, m);
      }
      if (k == 0) {
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
    {
      const int y = tid >> 4, x = tid & 15;
      if (x < tw && y < th) {                                // addBIOAvg4 (:109)
        const int blk = (y >> 2) * 4 + (x >> 2), i = (y + 1) * 18 + x + 1;
        const int b = sVxy[blk][0] * (sG[0][0][i] - sG[1][0][i]) + sVxy[blk][1] * (sG[0][1][i] - sG[1][1][i]);
        const int shiftNum = 15 - bd, offset = (1 << (shiftNum - 1)) + 2 * IFO;
        P.dst[0][(size_t)(by + y) * P.dstStride[0] + bx + x] = (int16_t)clip3(0, pmax, (int)(int16_t)((sP[0][i] + sP[1][i] + b + offset) >> shiftNum));
      }
    }
  }

  // ================================================================ chroma 4:2:0: 4-tap, both components
  if (!P.chroma) return;
  const int cw = tw >> 1, ch = th >> 1;
  const int nJobs = nList * 2;                               // (list, comp)
  for (int i = tid; i < nJobs * (cw + 3) * (ch + 3); i += 256) {
    const int job = i / ((cw + 3) * (ch + 3)), j = i - job * (cw + 3) * (ch + 3);
    const int l = bi ? (job >> 1) : l0, c = 1 + (job & 1);
    const int y = j / (cw + 3), x = j - y * (cw + 3);
    sCW[job][y * 12 + x] = (int16_t)ldw(R[l][c], win[l][1], org[l][1][0] + x - 1, org[l][1][1] + y - 1);
  }
  __syncthreads();
  for (int i = tid; i < nJobs * (ch + 3) * cw; i += 256) {
    const int job = i / ((ch + 3) * cw), j = i - job * (ch + 3) * cw;
    const int l = bi ? (job >> 1) : l0;
    const int y = j / cw, x = j - y * cw;
    const int8_t* f = kIfChroma + (mv[l][0] & 31) * 4;
    int s = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) s += f[t] * sCW[job][y * 12 + x + t];
    sCH[job][y * 8 + x] = (int16_t)((s - (IFO << sh1)) >> sh1);
  }
  __syncthreads();
  for (int i = tid; i < nJobs * ch * cw; i += 256) {
    const int job = i / (ch * cw), j = i - job * ch * cw;
    const int l = bi ? (job >> 1) : l0, c = 1 + (job & 1);
    const int y = j / cw, x = j - y * cw;
    const int8_t* f = kIfChroma + (mv[l][1] & 31) * 4;
    int s = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) s += f[t] * sCH[job][(y + t) * 8 + x];
    if (!bi) P.dst[c][(size_t)((by >> 1) + y) * P.dstStride[c] + (bx >> 1) + x] = (int16_t)clip3(0, pmax, (s + (1 << (5 + hr)) + (IFO << 6)) >> (6 + hr));
    else sCP[job][y * 8 + x] = (int16_t)(s >> 6);
  }
  if (bi) {
    __syncthreads();
    for (int i = tid; i < 2 * ch * cw; i += 256) {
      const int c = i / (ch * cw), j = i - c * ch * cw, y = j / cw, x = j - y * cw;
      P.dst[1 + c][(size_t)((by >> 1) + y) * P.dstStride[1 + c] + (bx >> 1) + x] =
          (int16_t)avg_bi(sCP[c][y * 8 + x], sCP[2 + c][y * 8 + x], dmvr ? 4 : pu.bcwW1, hr, pmax);
    }
  }
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

__global__ void __launch_bounds__(256) mc_affine_kernel(const McParams P)
{
  __shared__ int16_t sHf[2][16][9 * 4];     // per sub-block: 9 rows x 4 cols after horizontal filter
  __shared__ int16_t sE[2][16][36];         // PROF 6x6 buffers
  __shared__ int16_t sP[2][3][256];         // 14-bit predictions (luma 16x16, chroma 8x8 each)
  __shared__ int16_t sCH[2][2][4][7 * 4];   // chroma: [list][comp][sub-block] 7 rows x 4 cols

  const int tid = threadIdx.x;
  const uint32_t tile = P.tiles[blockIdx.x];
  const b200_pu pu = P.pus[tile >> 6];
  const int tx0 = (tile & 7) * 16, ty0 = ((tile >> 3) & 7) * 16;
  const int tw = min(16, pu.w - tx0), th = min(16, pu.h - ty0);
  const int bx = pu.x + tx0, by = pu.y + ty0;
  const int bd = P.bitDepth, pmax = (1 << bd) - 1, hr = max(2, 14 - bd), sh1 = 6 - hr;
  const bool bi = pu.refSlot[0] >= 0 && pu.refSlot[1] >= 0;
  const int nList = bi ? 2 : 1, l0 = pu.refSlot[0] >= 0 ? 0 : 1;
  const int hMin = (-P.ctuSize - 8 - pu.x + 1) * 16, hMax = (P.W + 8 - pu.x - 1) * 16;
  const int vMin = (-P.ctuSize - 8 - pu.y + 1) * 16, vMax = (P.H + 8 - pu.y - 1) * 16;

  AffModel M[2];
  RefPl R[2][3];
  for (int li = 0; li < nList; li++) {
    const int l = bi ? li : l0;
    aff_model(pu, l, M[l]);
    for (int c = 0; c < 3; c++) { R[l][c].p = P.refs[pu.refSlot[l] * 3 + c]; R[l][c].w = c ? P.W >> 1 : P.W; R[l][c].h = c ? P.H >> 1 : P.H; R[l][c].stride = P.refStride[c]; }
  }

  // ---- luma: sub-block sb = tid>>4 (4x4 grid in the tile), lane k = tid&15 ----
  const int sb = tid >> 4, k = tid & 15;
  const int sbx = (sb & 3) * 4, sby = (sb >> 2) * 4;
  const bool sbValid = sbx < tw && sby < th;
  for (int li = 0; li < nList; li++) {
    const int l = bi ? li : l0;
    int mx = 0, my = 0;
    if (sbValid) { aff_sub_mv(M[l], pu, (tx0 + sbx) >> 2, (ty0 + sby) >> 2, mx, my); mx = clip3(hMin, hMax, mx); my = clip3(vMin, vMax, my); }
    const int xF = mx & 15, yF = my & 15, X0 = bx + sbx + (mx >> 4), Y0 = by + sby + (my >> 4);
    if (sbValid) {
      const int8_t* fh = kIfLuma4x4 + xF * 8;
      for (int j = k; j < 36; j += 16) {                      // 9 rows (y-2..y+6: 6-tap taps 1..6 of the 8-tap array) x 4 cols
        const int y = j >> 2, x = j & 3;
        int s = 0;
        if (xF == 0) s = 64 * ldc(R[l][0], X0 + x, Y0 + y - 2);
        else {
#pragma unroll
          for (int t = 1; t < 7; t++) s += fh[t] * ldc(R[l][0], X0 + x + t - 3, Y0 + y - 2);
        }
        sHf[l][sb][j] = (int16_t)((s - (IFO << sh1)) >> sh1);
      }
      if (M[l].prof) {                                       // ring of the 6x6 PROF buffer from integer samples (:1233-1262)
        const int rx = X0 + (xF >> 3) - 1, ry = Y0 + (yF >> 3) - 1;
        for (int j = k; j < 36; j += 16) {
          const int y = j / 6, x = j - y * 6;
          if (x > 0 && x < 5 && y > 0 && y < 5) continue;
          sE[l][sb][j] = (int16_t)((int16_t)(ldc(R[l][0], rx + x, ry + y) << hr) - IFO);
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
      if (M[l].prof) sE[l][sb][(y + 1) * 6 + x + 1] = (int16_t)(s >> 6);
      else if (bi)   sP[l][0][(sby + y) * 16 + sbx + x] = (int16_t)(s >> 6);
      else P.dst[0][(size_t)(by + sby + y) * P.dstStride[0] + bx + sbx + x] = (int16_t)clip3(0, pmax, (s + (1 << (5 + hr)) + (IFO << 6)) >> (6 + hr));
    }
    __syncwarp();
    if (sbValid && M[l].prof) {                              // gradFilterCore<false> :212 + applyPROFCore :61
      const int y = k >> 2, x = k & 3, c = (y + 1) * 6 + x + 1;
      const int16_t* E = sE[l][sb];
      const int gX = (E[c + 1] >> 6) - (E[c - 1] >> 6), gY = (E[c + 6] >> 6) - (E[c - 6] >> 6);
      // dMv of sample (x,y) inside the 4x4 (:1043-1090): linear in x,y, then rounded by 8 and clipped to +-31
      const int qHX = M[l].dHX * 4, qHY = M[l].dHY * 4, qVX = M[l].dVX * 4, qVY = M[l].dVY * 4;
      int dh = ((M[l].dHX + M[l].dVX) * 2) - ((qHX + qVX) * 2) + x * qHX + y * qVX;
      int dv = ((M[l].dHY + M[l].dVY) * 2) - ((qHY + qVY) * 2) + x * qHY + y * qVY;
      round_affine(dh, dv, 8);
      dh = clip3(-31, 31, dh); dv = clip3(-31, 31, dv);
      const int lim = 1 << max(bd + 1, 13);
      const int dI = clip3(-lim, lim - 1, dh * gX + dv * gY);
      int v = (int16_t)(E[c] + dI);
      if (bi) sP[l][0][(sby + y) * 16 + sbx + x] = (int16_t)v;
      else P.dst[0][(size_t)(by + sby + y) * P.dstStride[0] + bx + sbx + x] = (int16_t)clip3(0, pmax, (int)(int16_t)((v + (1 << (hr - 1)) + IFO) >> hr));
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
    if (valid) {
      int ax, ay, bxm, bym;
      const int i0 = (tx0 >> 2) + (csx >> 1), j0 = (ty0 >> 2) + (csy >> 1);
      aff_sub_mv(M[l], pu, i0, j0, ax, ay); aff_sub_mv(M[l], pu, i0 + 1, j0 + 1, bxm, bym);
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
        for (int t = 0; t < 4; t++) s += fh[t] * ldc(R[l][c], X0 + x + t - 1, Y0 + y - 1);
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
      else P.dst[c][(size_t)((by >> 1) + csy + y) * P.dstStride[c] + (bx >> 1) + csx + x] = (int16_t)clip3(0, pmax, (s + (1 << (5 + hr)) + (IFO << 6)) >> (6 + hr));
    }
  }
  if (!bi) return;
  __syncthreads();
  {
    const int y = tid >> 4, x = tid & 15;
    if (x < tw && y < th) P.dst[0][(size_t)(by + y) * P.dstStride[0] + bx + x] = (int16_t)avg_bi(sP[0][0][y * 16 + x], sP[1][0][y * 16 + x], pu.bcwW1, hr, pmax);
    if (P.chroma && tid < 128) {
      const int c = tid >> 6, j = tid & 63, yy = j >> 3, xx = j & 7;
      if (xx < (tw >> 1) && yy < (th >> 1))
        P.dst[1 + c][(size_t)((by >> 1) + yy) * P.dstStride[1 + c] + (bx >> 1) + xx] = (int16_t)avg_bi(sP[0][1 + c][yy * 8 + xx], sP[1][1 + c][yy * 8 + xx], pu.bcwW1, hr, pmax);
    }
  }
}

int launch_mc(const McLaunch& L, cudaStream_t s, KProf* prof)
{
  McParams P;
  for (int c = 0; c < 3; c++) { P.dst[c] = L.dst.p[c]; P.dstStride[c] = L.dst.stride[c]; P.refStride[c] = L.refStride[c]; }
  for (int i = 0; i < B200_MAX_SLOTS * 3; i++) P.refs[i] = L.refs[i];
  P.W = L.geom.width; P.H = L.geom.height; P.bitDepth = L.geom.bitDepth; P.ctuSize = L.geom.ctuSize; P.chroma = L.geom.chromaFormat == 1;
  P.pus = L.pus; P.dmvrMv = L.dmvrMv;
  if (L.numTilesT) { if (prof) prof->begin(B200_KF_MC_TILE, s); P.tiles = L.tilesT; P.numTiles = L.numTilesT; mc_tile_kernel<<<L.numTilesT, 256, 0, s>>>(P); B200_CUDA(cudaGetLastError()); if (prof) prof->end(B200_KF_MC_TILE, s); }
  if (L.numTilesA) { if (prof) prof->begin(B200_KF_MC_AFFINE, s); P.tiles = L.tilesA; P.numTiles = L.numTilesA; mc_affine_kernel<<<L.numTilesA, 256, 0, s>>>(P); B200_CUDA(cudaGetLastError()); if (prof) prof->end(B200_KF_MC_AFFINE, s); }
  return 0;
}

}  // namespace b200
