/* oracle/k3_deblock.c — CPU restatement of K3 (VVC deblocking on the 4x4 luma / 8x8 chroma grid).
 * TEST INFRASTRUCTURE ONLY — see vvc_oracle.h. Pinned against oracle/_ref (tests/test_k3_oracle_vs_ref.py). */
#include "vvc_oracle.h"
#include <stdlib.h>

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* LoopFilter.cpp:84-92 */
static const uint16_t tcTable[66] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,3,4,4,4,4,5,5,5,5,7,7,8,9,10,10,11,13,14,15,17,19,21,24,25,29,33,36,
  41,45,51,57,64,71,80,89,100,112,125,141,157,177,198,222,250,280,314,352,395 };
static const uint8_t betaTable[64] = { 0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,
  44,46,48,50,52,54,56,58,60,62,64,66,68,70,72,74,76,78,80,82,84,86,88 };

#define P(i) src[-(ptrdiff_t)((i) + 1) * offset]   /* p_i */
#define Q(i) src[(ptrdiff_t)(i) * offset]          /* q_i */

/* LoopFilter.cpp:213-279 */
static void pel_filter_luma_line(int16_t* src, ptrdiff_t offset, int tc, int sw, int thrCut, int fsP, int fsQ, int bd)
{
  const int pmax = (1 << bd) - 1;
  const int m1 = P(2), m2 = P(1), m3 = P(0), m4 = Q(0), m5 = Q(1), m6 = Q(2);
  if (sw) {
    const int m0 = P(3), m7 = Q(3);
    P(2) = (int16_t)clip3(m1 - 1 * tc, m1 + 1 * tc, (2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3);
    P(1) = (int16_t)clip3(m2 - 2 * tc, m2 + 2 * tc, (m1 + m2 + m3 + m4 + 2) >> 2);
    P(0) = (int16_t)clip3(m3 - 3 * tc, m3 + 3 * tc, (m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3);
    Q(0) = (int16_t)clip3(m4 - 3 * tc, m4 + 3 * tc, (m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3);
    Q(1) = (int16_t)clip3(m5 - 2 * tc, m5 + 2 * tc, (m3 + m4 + m5 + m6 + 2) >> 2);
    Q(2) = (int16_t)clip3(m6 - 1 * tc, m6 + 1 * tc, (m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3);
  } else {
    int delta = (9 * (m4 - m3) - 3 * (m5 - m2) + 8) >> 4;
    if (iabs(delta) < thrCut) {
      delta = clip3(-tc, tc, delta);
      const int tc2 = tc >> 1;
      P(0) = (int16_t)clip3(0, pmax, m3 + delta);
      if (fsP) P(1) = (int16_t)clip3(0, pmax, m2 + clip3(-tc2, tc2, ((((m1 + m3 + 1) >> 1) - m2 + delta) >> 1)));
      Q(0) = (int16_t)clip3(0, pmax, m4 - delta);
      if (fsQ) Q(1) = (int16_t)clip3(0, pmax, m5 + clip3(-tc2, tc2, ((((m6 + m4 + 1) >> 1) - m5 - delta) >> 1)));
    }
  }
}

void orc_lf_pel_filter_luma(int16_t* src, ptrdiff_t step, ptrdiff_t offset, int tc, int sw, int thrCut, int fsP, int fsQ, int bd)
{
  for (int i = 0; i < 4; i++) pel_filter_luma_line(src + step * i, offset, tc, sw, thrCut, fsP, fsQ, bd);
}

/* LoopFilter.cpp:102-196: long (bilinear) luma filters */
void orc_lf_filtering_pq(int16_t* base, ptrdiff_t step, ptrdiff_t offset, int nP, int nQ, int tc)
{
  static const int c7[7] = { 59, 50, 41, 32, 23, 14, 5 }, c5[5] = { 58, 45, 32, 19, 6 }, c3[3] = { 53, 32, 11 };
  static const int t7[7] = { 6, 5, 4, 3, 2, 1, 1 }, t3[3] = { 6, 4, 2 };
  const int* cP = nP == 7 ? c7 : nP == 5 ? c5 : c3;
  const int* cQ = nQ == 7 ? c7 : nQ == 5 ? c5 : c3;
  const int* tP = nP == 3 ? t3 : t7;
  const int* tQ = nQ == 3 ? t3 : t7;
  for (int l = 0; l < 4; l++) {
    int16_t* src = base + step * l;
    const int refP = (P(nP - 1) + P(nP) + 1) >> 1;
    const int refQ = (Q(nQ - 1) + Q(nQ) + 1) >> 1;
    int mid;
    if (nP == nQ) {
      if (nP == 5) mid = (2 * (P(0) + Q(0) + P(1) + Q(1) + P(2) + Q(2)) + P(3) + Q(3) + P(4) + Q(4) + 8) >> 4;
      else         mid = (2 * (P(0) + Q(0)) + P(1) + Q(1) + P(2) + Q(2) + P(3) + Q(3) + P(4) + Q(4) + P(5) + Q(5) + P(6) + Q(6) + 8) >> 4;
    } else {
      const int big = nP > nQ ? nP : nQ, small = nP > nQ ? nQ : nP;
      if (big == 7 && small == 5) mid = (2 * (P(0) + Q(0) + P(1) + Q(1)) + P(2) + Q(2) + P(3) + Q(3) + P(4) + Q(4) + P(5) + Q(5) + 8) >> 4;
      else if (big == 7 && small == 3) {
        /* L = long side samples, S = short side samples, index 0 nearest the edge */
        #define L(i) (nP > nQ ? P(i) : Q(i))
        #define S(i) (nP > nQ ? Q(i) : P(i))
        mid = (2 * (L(0) + S(0)) + S(0) + 2 * (S(1) + S(2)) + L(1) + S(1) + L(2) + L(3) + L(4) + L(5) + L(6) + 8) >> 4;
        #undef L
        #undef S
      } else mid = (P(0) + Q(0) + P(1) + Q(1) + P(2) + Q(2) + P(3) + Q(3) + 4) >> 3;
    }
    for (int i = 0; i < nP; i++) { const int s = P(i), cv = (tc * tP[i]) >> 1; P(i) = (int16_t)clip3(s - cv, s + cv, (mid * cP[i] + refP * (64 - cP[i]) + 32) >> 6); }
    for (int i = 0; i < nQ; i++) { const int s = Q(i), cv = (tc * tQ[i]) >> 1; Q(i) = (int16_t)clip3(s - cv, s + cv, (mid * cQ[i] + refQ * (64 - cQ[i]) + 32) >> 6); }
  }
}

/* LoopFilter.cpp:1410-1461 */
static int use_strong(const int16_t* src, ptrdiff_t offset, int d, int beta, int tc, int largeP, int largeQ, int maxP, int maxQ, int chromaHorCtb)
{
  const int m3 = P(0), m4 = Q(0);
  if (!(d < (beta >> 2) && iabs(m3 - m4) < ((tc * 5 + 1) >> 1))) return 0;
  const int m0 = P(3), m7 = Q(3), m2 = P(1);
  int sp3 = chromaHorCtb ? iabs(m2 - m3) : iabs(m0 - m3);
  int sq3 = iabs(m7 - m4);
  if (largeP || largeQ) {
    if (largeP) {
      const int mP4 = P(maxP);
      if (maxP == 7) sp3 += iabs(P(4) - P(5) - P(6) + mP4);
      sp3 = (sp3 + iabs(m0 - mP4) + 1) >> 1;
    }
    if (largeQ) {
      const int m11 = Q(maxQ);
      if (maxQ == 7) sq3 += iabs(Q(4) - Q(5) - Q(6) + m11);
      sq3 = (sq3 + iabs(m11 - m7) + 1) >> 1;
    }
    return (sp3 + sq3) < (beta * 3 >> 5) && d < (beta >> 4) && iabs(m3 - m4) < ((tc * 5 + 1) >> 1);
  }
  return (sp3 + sq3) < (beta >> 3);
}

static inline int calc_dp(const int16_t* src, ptrdiff_t offset) { return iabs(P(2) - 2 * P(1) + P(0)); }
static inline int calc_dp_ctb(const int16_t* src, ptrdiff_t offset) { return iabs(P(1) - 2 * P(1) + P(0)); }   /* LoopFilter.cpp:1395 */
static inline int calc_dq(const int16_t* src, ptrdiff_t offset) { return iabs(Q(0) - 2 * Q(1) + Q(2)); }

static int tc_value(int idx, int bd) { return bd < 10 ? (tcTable[idx] + (1 << (9 - bd))) >> (10 - bd) : tcTable[idx] << (bd - 10); }

/* LoopFilter.cpp:1463-1617 xEdgeFilterLuma for one 4-sample segment whose first Q sample is `src`. */
static void edge_luma(int16_t* src, ptrdiff_t offset, ptrdiff_t step, const b200_lf_param* lfp, const b200_lf_slice* sl,
                      const b200_lf_seq* seq, int bd, int horCtbBoundary, int isVer, ptrdiff_t stride)
{
  const int bs = lfp->bs & 3;
  if (!bs) return;
  int qp = lfp->qp[0];
  if (seq && seq->ladfEnabled) {     /* LoopFilter.cpp:1363 deriveLADFShift */
    int shift = seq->ladfQpOffset[0];
    const int lvl = isVer ? (src[0] + src[3 * stride] + src[-1] + src[3 * stride - 1]) >> 2
                          : (src[0] + src[3] + src[-stride] + src[-stride + 3]) >> 2;
    for (int k = 1; k < seq->ladfNumIntervals; k++) { if (lvl > seq->ladfIntervalLowerBound[k]) shift = seq->ladfQpOffset[k]; else break; }
    qp += shift;
  }
  const int maxP = (lfp->sideMaxFiltLength >> 4) & 7, maxQ = lfp->sideMaxFiltLength & 7;
  int largeP = maxP > 3; const int largeQ = maxQ > 3;
  if (horCtbBoundary) largeP = 0;
  const int idxTc = clip3(0, 65, qp + 2 * (bs - 1) + sl->tcOffsetDiv2[0] * 2);
  const int idxB  = clip3(0, 63, qp + sl->betaOffsetDiv2[0] * 2);
  const int tc = tc_value(idxTc, bd), beta = betaTable[idxB] << (bd - 8);
  const int sideThr = (beta + (beta >> 1)) >> 3, thrCut = tc * 10;
  const int16_t* s0 = src; const int16_t* s3 = src + 3 * step;
  const int dp0 = calc_dp(s0, offset), dq0 = calc_dq(s0, offset), dp3 = calc_dp(s3, offset), dq3 = calc_dq(s3, offset);
  const int d0 = dp0 + dq0, d3 = dp3 + dq3;
  if (largeP || largeQ) {
    const ptrdiff_t o3 = 3 * offset;
    const int dp0L = largeP ? (dp0 + calc_dp(s0 - o3, offset) + 1) >> 1 : dp0;
    const int dq0L = largeQ ? (dq0 + calc_dq(s0 + o3, offset) + 1) >> 1 : dq0;
    const int dp3L = largeP ? (dp3 + calc_dp(s3 - o3, offset) + 1) >> 1 : dp3;
    const int dq3L = largeQ ? (dq3 + calc_dq(s3 + o3, offset) + 1) >> 1 : dq3;
    const int d0L = dp0L + dq0L, d3L = dp3L + dq3L;
    if (d0L + d3L < beta) {
      if (use_strong(s0, offset, 2 * d0L, beta, tc, largeP, largeQ, maxP, maxQ, 0) &&
          use_strong(s3, offset, 2 * d3L, beta, tc, largeP, largeQ, maxP, maxQ, 0)) {
        orc_lf_filtering_pq(src, step, offset, largeP ? maxP : 3, largeQ ? maxQ : 3, tc);
        return;
      }
    }
  }
  if (d0 + d3 < beta) {
    int fP = 0, fQ = 0, sw = 0;
    if (maxP > 1 && maxQ > 1) { fP = (dp0 + dp3) < sideThr; fQ = (dq0 + dq3) < sideThr; }
    if (maxP > 2 && maxQ > 2) sw = use_strong(s0, offset, 2 * d0, beta, tc, 0, 0, 7, 7, 0) && use_strong(s3, offset, 2 * d3, beta, tc, 0, 0, 7, 7, 0);
    orc_lf_pel_filter_luma(src, step, offset, tc, sw, thrCut, fP, fQ, bd);
  }
}

/* LoopFilter.cpp:281-332 */
static void pel_filter_chroma(int16_t* src, ptrdiff_t offset, int tc, int sw, int bd, int horCtb)
{
  const int pmax = (1 << bd) - 1;
  const int m2 = P(1), m3 = P(0), m4 = Q(0), m5 = Q(1);
  if (sw) {
    const int m6 = Q(2), m7 = Q(3);
    if (horCtb) {
      P(0) = (int16_t)clip3(m3 - tc, m3 + tc, (3 * m2 + 2 * m3 + m4 + m5 + m6 + 4) >> 3);
      Q(0) = (int16_t)clip3(m4 - tc, m4 + tc, (2 * m2 + m3 + 2 * m4 + m5 + m6 + m7 + 4) >> 3);
      Q(1) = (int16_t)clip3(m5 - tc, m5 + tc, (m2 + m3 + m4 + 2 * m5 + m6 + 2 * m7 + 4) >> 3);
      Q(2) = (int16_t)clip3(m6 - tc, m6 + tc, (m3 + m4 + m5 + 2 * m6 + 3 * m7 + 4) >> 3);
    } else {
      const int m0 = P(3), m1 = P(2);
      P(2) = (int16_t)clip3(m1 - tc, m1 + tc, (3 * m0 + 2 * m1 + m2 + m3 + m4 + 4) >> 3);
      P(1) = (int16_t)clip3(m2 - tc, m2 + tc, (2 * m0 + m1 + 2 * m2 + m3 + m4 + m5 + 4) >> 3);
      P(0) = (int16_t)clip3(m3 - tc, m3 + tc, (m0 + m1 + m2 + 2 * m3 + m4 + m5 + m6 + 4) >> 3);
      Q(0) = (int16_t)clip3(m4 - tc, m4 + tc, (m1 + m2 + m3 + 2 * m4 + m5 + m6 + m7 + 4) >> 3);
      Q(1) = (int16_t)clip3(m5 - tc, m5 + tc, (m2 + m3 + m4 + 2 * m5 + m6 + 2 * m7 + 4) >> 3);
      Q(2) = (int16_t)clip3(m6 - tc, m6 + tc, (m3 + m4 + m5 + 2 * m6 + 3 * m7 + 4) >> 3);
    }
  } else {
    const int delta = clip3(-tc, tc, (((m4 - m3) * 4) + m2 - m5 + 4) >> 3);
    P(0) = (int16_t)clip3(0, pmax, m3 + delta);
    Q(0) = (int16_t)clip3(0, pmax, m4 - delta);
  }
}

/* LoopFilter.cpp:1619-1731 xEdgeFilterChroma for one component, one 2-sample segment (4:2:0). */
static void edge_chroma(int16_t* src, ptrdiff_t offset, ptrdiff_t step, const b200_lf_param* lfp, const b200_lf_slice* sl,
                        int c /*1|2*/, int bd, int horCtb)
{
  const int bs = (lfp->bs >> (2 * c)) & 3;
  const int large = (lfp->flags >> 5) & 1;
  if (!(bs == 2 || (large && bs == 1))) return;
  const int qp = lfp->qp[c];
  const int idxTc = clip3(0, 65, qp + 2 * (bs - 1) + sl->tcOffsetDiv2[c] * 2);
  const int tc = tc_value(idxTc, bd);
  if (large) {
    const int idxB = clip3(0, 63, qp + sl->betaOffsetDiv2[c] * 2);
    const int beta = betaTable[idxB] * (1 << (bd - 8));
    const int16_t* s1 = src + step;    /* subSamplingShift == 1 (4:2:0): second decision line is line 1 */
    const int dp0 = horCtb ? calc_dp_ctb(src, offset) : calc_dp(src, offset), dq0 = calc_dq(src, offset);
    const int dp3 = horCtb ? calc_dp_ctb(s1, offset) : calc_dp(s1, offset), dq3 = calc_dq(s1, offset);
    const int d0 = dp0 + dq0, d3 = dp3 + dq3;
    if (d0 + d3 < beta) {
      const int sw = use_strong(src, offset, 2 * d0, beta, tc, 0, 0, 7, 7, horCtb) && use_strong(s1, offset, 2 * d3, beta, tc, 0, 0, 7, 7, horCtb);
      for (int i = 0; i < 2; i++) pel_filter_chroma(src + step * i, offset, tc, sw, bd, horCtb);
      return;
    }
  }
  for (int i = 0; i < 2; i++) pel_filter_chroma(src + step * i, offset, tc, 0, bd, horCtb);
}

void orc_lf_deblock(const b200_geom* g, int16_t* const planes[3], const b200_lf_param* lfV, const b200_lf_param* lfH,
                    const uint8_t* ctuSlice, const b200_lf_slice* slices, const b200_lf_seq* seq, int dirs)
{
  const int W = g->width, H = g->height, W4 = (W + 3) >> 2, H4 = (H + 3) >> 2, bd = g->bitDepth;
  const int ctuLog2 = g->ctuSize == 128 ? 7 : g->ctuSize == 64 ? 6 : 5;
  const int ctusW = (W + g->ctuSize - 1) >> ctuLog2;
  for (int dir = 0; dir < 2; dir++) {
    if (!(dirs & (1 << dir))) continue;
    const b200_lf_param* grid = dir ? lfH : lfV;
    for (int y4 = 0; y4 < H4; y4++)
      for (int x4 = 0; x4 < W4; x4++) {
        const b200_lf_param* lfp = &grid[y4 * W4 + x4];
        const int x = x4 * 4, y = y4 * 4;
        const b200_lf_slice* sl = &slices[ctuSlice ? ctuSlice[(y >> ctuLog2) * ctusW + (x >> ctuLog2)] : 0];
        if (sl->disable) continue;
        if (lfp->bs & 3) {
          int16_t* src = planes[0] + (size_t)y * g->stride[0] + x;
          if (dir == 0) edge_luma(src, 1, g->stride[0], lfp, sl, seq, bd, 0, 1, g->stride[0]);
          else          edge_luma(src, g->stride[0], 1, lfp, sl, seq, bd, (y & (g->ctuSize - 1)) == 0, 0, g->stride[0]);
        }
        /* chroma 4:2:0: edges on the 8-sample chroma grid = every 16 luma samples across the edge direction */
        if (g->chromaFormat == 1 && (lfp->bs >> 2) && ((dir == 0 ? x : y) & 15) == 0) {
          const int cx = x >> 1, cy = y >> 1;
          const int horCtb = dir == 1 && (cy & ((g->ctuSize >> 1) - 1)) == 0;
          for (int c = 1; c <= 2; c++) {
            int16_t* src = planes[c] + (size_t)cy * g->stride[c] + cx;
            if (dir == 0) edge_chroma(src, 1, g->stride[c], lfp, sl, c, bd, 0);
            else          edge_chroma(src, g->stride[c], 1, lfp, sl, c, bd, horCtb);
          }
        }
      }
  }
}
