/* oracle/k2_inter.c — CPU restatement of K2 (inter prediction: 8/4-tap MC, bi-pred average / BCW, BDOF, DMVR, affine + PROF).
 * TEST INFRASTRUCTURE ONLY — see vvc_oracle.h. Pinned against oracle/_ref (tests/test_k2_oracle_vs_ref.py):
 * the real InterPrediction::motionCompensation runs on real CodingUnits / reference Pictures built by the shim. */
#include "vvc_oracle.h"
#include "../vvdec_b200/csrc/vvc_tables.h"
#include <string.h>
#include <stdlib.h>

#define IF_OFFS 8192            /* IF_INTERNAL_OFFS = 1 << (IF_INTERNAL_PREC-1), CommonDef.h */
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

typedef struct { const int16_t* p; int w, h, stride; } Plane;
static inline int px(const Plane* r, int x, int y) { return r->p[(size_t)clip3(0, r->h - 1, y) * r->stride + clip3(0, r->w - 1, x)]; }

/* Mv.cpp:64 clipMvInPic */
static void clip_mv(int mv[2], int x, int y, const b200_geom* g)
{
  const int hMax = (g->width + 8 - x - 1) * 16, hMin = (-g->ctuSize - 8 - x + 1) * 16;
  const int vMax = (g->height + 8 - y - 1) * 16, vMin = (-g->ctuSize - 8 - y + 1) * 16;
  mv[0] = clip3(hMin, hMax, mv[0]); mv[1] = clip3(vMin, vMax, mv[1]);
}

/* Separable N-tap interpolation of a w x h block whose top-left integer reference position is (x0,y0).
 * InterpolationFilter.cpp:556 filter<> chained as xPredInterBlk (:750-870) does; a zero fraction is the {..,64,..} tap set, for which
 * the one-stage and copy paths of the reference (filterCopy :424, single filterHor/Ver) give the same integers.
 * last != 0: rounded/clipped samples (uni-prediction); last == 0: 14-bit intermediates biased by -IF_OFFS. */
static void interp(const Plane* ref, int x0, int y0, const int8_t* fh, const int8_t* fv, int taps, int w, int h, int bd, int last,
                   int16_t* out, int ostride)
{
  const int half = taps / 2 - 1, headroom = 14 - bd < 2 ? 2 : 14 - bd, pmax = (1 << bd) - 1;
  const int sh1 = 6 - headroom;
  for (int x = 0; x < w; x++) {
    int col[128 + 8];
    for (int y = 0; y < h + taps - 1; y++) {
      int s = 0;
      for (int t = 0; t < taps; t++) s += fh[t] * px(ref, x0 + x + t - half, y0 + y - half);
      col[y] = (int16_t)((s - (IF_OFFS << sh1)) >> sh1);
    }
    for (int y = 0; y < h; y++) {
      int s = 0;
      for (int t = 0; t < taps; t++) s += fv[t] * col[y + t];
      out[y * ostride + x] = last ? (int16_t)clip3(0, pmax, (s + (1 << (5 + headroom)) + (IF_OFFS << 6)) >> (6 + headroom)) : (int16_t)(s >> 6);
    }
  }
}

static const int8_t k64_8[8] = { 0, 0, 0, 64, 0, 0, 0, 0 };

/* luma tap selection of InterpolationFilter::filterHor/Ver (:1044-1212) and filter4x4/8xH/16xH (:669-733) */
static const int8_t* luma_taps(int frac, int w, int h, int altHpel)
{
  if (frac == 0) return k64_8;
  if (w == 4 && h == 4) return &kIfLuma4x4[frac * 8];
  if (frac == 8 && altHpel) return kIfAltHpel;
  return &kIfLuma[frac * 8];
}

/* xPredInterBlk (:750) for one component of a block at luma (bx,by) size (bw,bh) with an already clipped mv. */
static void pred_block(const b200_geom* g, const Plane* ref, int comp, int bx, int by, int bw, int bh, const int mv[2], int altHpel, int last,
                       int16_t* out, int ostride)
{
  if (comp == 0) {
    const int xF = mv[0] & 15, yF = mv[1] & 15;
    interp(ref, bx + (mv[0] >> 4), by + (mv[1] >> 4), luma_taps(xF, bw, bh, altHpel), luma_taps(yF, bw, bh, altHpel), 8, bw, bh, g->bitDepth, last, out, ostride);
  } else {
    const int xF = mv[0] & 31, yF = mv[1] & 31;
    interp(ref, (bx >> 1) + (mv[0] >> 5), (by >> 1) + (mv[1] >> 5), &kIfChroma[xF * 4], &kIfChroma[yF * 4], 4, bw >> 1, bh >> 1, g->bitDepth, last, out, ostride);
  }
}

/* xWeightedAverage (:1346): addAvg (Buffer.cpp:441 / :67) or BCW addWeightedAvg (Buffer.cpp:372) */
static void average(const int16_t* a, const int16_t* b, int w, int h, int sstride, int bd, int w1, int16_t* d, int ds)
{
  const int pmax = (1 << bd) - 1, hr = 14 - bd < 2 ? 2 : 14 - bd;
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    const int p0 = a[y * sstride + x], p1 = b[y * sstride + x];
    int v;
    if (w1 == 4) v = (p0 + p1 + (1 << hr) + 2 * IF_OFFS) >> (hr + 1);
    else         v = (p0 * (8 - w1) + p1 * w1 + (1 << (hr + 2)) + (IF_OFFS << 3)) >> (hr + 3);
    d[y * ds + x] = (int16_t)clip3(0, pmax, v);
  }
}

/* explicit weighted prediction: addWeightBi (WeightPrediction.cpp:164-236, via wghtAvg = addWeightedAvgCore) / addWeightUni (:238-331);
 * a, b: 14-bit intermediates (b == NULL: uni-prediction) */
static void weighted(const int16_t* a, const int16_t* b, int w, int h, int sstride, int bd, const b200_wp* e, int comp, int16_t* d, int ds)
{
  const int pmax = (1 << bd) - 1, shiftNum = 14 - bd < 2 ? 2 : 14 - bd, shift = e->shift[comp] + shiftNum;
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    int v;
    if (b) v = (e->w0[comp] * (a[y * sstride + x] + IF_OFFS) + e->w1[comp] * (b[y * sstride + x] + IF_OFFS) + ((1 << shift) >> 1) + e->offset[comp] * (1 << (shift - 1))) >> shift;
    else   v = ((e->w0[comp] * (a[y * sstride + x] + IF_OFFS) + (shift > 0 ? 1 << (shift - 1) : 0)) >> shift) + e->offset[comp];
    d[y * ds + x] = (int16_t)clip3(0, pmax, v);
  }
}

/* GEO blending (InterpolationFilter.cpp:1217 xWeightedGeoBlk): a = partition 0 (tmpGeoBuf0), b = partition 1, 14-bit intermediates */
static void geo_blend(const int16_t* a, const int16_t* b, int w, int h, int sstride, int bd, int comp, int cuW, int cuH, int splitDir, int16_t* d, int ds)
{
  const int pmax = (1 << bd) - 1, shift = (14 - bd < 2 ? 2 : 14 - bd) + 3, offset = (1 << (shift - 1)) + (IF_OFFS << 3);
  const int sc = comp ? 1 : 0, angle = kGeoParams[splitDir * 2];
  int l2w = 0, l2h = 0; while ((1 << l2w) < cuW) l2w++; while ((1 << l2h) < cuH) l2h++;
  const int16_t* wo = &kGeoWeightOffset[((splitDir * 4 + (l2h - 3)) * 4 + (l2w - 3)) * 2];
  const uint8_t* M = &kGeoWeights[(size_t)kGeoAngle2Mask[angle] * VVC_GEO_MASK_SIZE * VVC_GEO_MASK_SIZE];
  const int mir = kGeoAngle2Mirror[angle];
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    const int row = mir == 2 ? VVC_GEO_MASK_SIZE - 1 - wo[1] - (y << sc) : wo[1] + (y << sc);
    const int col = mir == 1 ? VVC_GEO_MASK_SIZE - 1 - wo[0] - (x << sc) : wo[0] + (x << sc);
    const int wt = M[row * VVC_GEO_MASK_SIZE + col];
    const int v = (wt * a[y * sstride + x] + (8 - wt) * b[y * sstride + x] + offset) >> shift;
    d[y * ds + x] = (int16_t)clip3(0, pmax, v);
  }
}

/* ---- BDOF: xPredInterBlk bio tail (:847-885) + applyBiOptFlow (:1290) + gradFilterCore<true> (:212) + BiOptFlowCore (:162) ---- */
static inline int shift_msb(int numer, int denom) { int m = 0; while (m < 32 && denom >= (1 << m)) m++; return numer >> (m - 1); }   /* rightShiftMSB :92 */

/* P: (w+2)x(h+2) 14-bit predictions incl. the 1-sample ring. The ring comes from integer reference samples (no interpolation). */
static void bdof_fill(const b200_geom* g, const Plane* ref, int bx, int by, int w, int h, const int mv[2], int altHpel, int16_t* P)
{
  const int S = w + 2, shift = 14 - g->bitDepth < 2 ? 2 : 14 - g->bitDepth;
  pred_block(g, ref, 0, bx, by, w, h, mv, altHpel, 0, P + S + 1, S);
  const int xF = mv[0] & 15, yF = mv[1] & 15;
  const int rx = bx + (mv[0] >> 4) - (xF < 8 ? 1 : 0), ry = by + (mv[1] >> 4) - (yF < 8 ? 1 : 0);     /* ring origin = sample (-1,-1) of P */
  for (int y = 0; y < h + 2; y++) for (int x = 0; x < w + 2; x++) {
    if (x > 0 && x < w + 1 && y > 0 && y < h + 1) continue;
    P[y * S + x] = (int16_t)((int16_t)(px(ref, rx + x, ry + y) << shift) - IF_OFFS);
  }
}

static void bdof(const int16_t* P0, const int16_t* P1, int w, int h, int bd, int16_t* dst, int ds)
{
  const int S = w + 2, pmax = (1 << bd) - 1;
  /* gradients on the interior, replicated into the ring; the prediction ring itself is then overwritten by replication
     (gradFilterCore<true> pads src too, :236-266) */
  int16_t *gx[2], *gy[2], *Q[2];
  for (int l = 0; l < 2; l++) {
    const int16_t* P = l ? P1 : P0;
    gx[l] = (int16_t*)calloc((size_t)S * (h + 2), 2); gy[l] = (int16_t*)calloc((size_t)S * (h + 2), 2); Q[l] = (int16_t*)malloc((size_t)S * (h + 2) * 2);
    memcpy(Q[l], P, (size_t)S * (h + 2) * 2);
    for (int y = 1; y <= h; y++) for (int x = 1; x <= w; x++) {
      gx[l][y * S + x] = (int16_t)((P[y * S + x + 1] >> 6) - (P[y * S + x - 1] >> 6));
      gy[l][y * S + x] = (int16_t)((P[(y + 1) * S + x] >> 6) - (P[(y - 1) * S + x] >> 6));
    }
    int16_t* arr[3] = { gx[l], gy[l], Q[l] };
    for (int k = 0; k < 3; k++) {
      int16_t* A = arr[k];
      for (int y = 1; y <= h; y++) { A[y * S] = A[y * S + 1]; A[y * S + w + 1] = A[y * S + w]; }
      memcpy(A, A + S, S * 2); memcpy(A + (h + 1) * S, A + h * S, S * 2);
    }
  }
  const int shiftNum = 14 + 1 - bd, offset = (1 << (shiftNum - 1)) + 2 * IF_OFFS, limit = 15;
  for (int by = 0; by < h; by += 4) for (int bx = 0; bx < w; bx += 4) {
    int sAbsGX = 0, sAbsGY = 0, sDIX = 0, sDIY = 0, sSign = 0;
    for (int y = 0; y < 6; y++) for (int x = 0; x < 6; x++) {          /* calcBIOSums :134: 6x6 window = 4x4 block + ring */
      const int i = (by + y) * S + bx + x;
      const int tGX = (gx[0][i] + gx[1][i]) >> 1, tGY = (gy[0][i] + gy[1][i]) >> 1;
      const int tDI = (Q[1][i] >> 4) - (Q[0][i] >> 4);
      sAbsGX += iabs(tGX); sAbsGY += iabs(tGY);
      sDIX += tGX < 0 ? -tDI : (tGX == 0 ? 0 : tDI);
      sDIY += tGY < 0 ? -tDI : (tGY == 0 ? 0 : tDI);
      sSign += tGY < 0 ? -tGX : (tGY == 0 ? 0 : tGX);
    }
    int tmpx = sAbsGX == 0 ? 0 : shift_msb(sDIX * 4, sAbsGX);
    tmpx = clip3(-limit, limit, tmpx);
    const int mainG = sSign >> 12, secG = sSign & ((1 << 12) - 1);
    int tmpData = tmpx * mainG;
    tmpData = ((tmpData * (1 << 12)) + tmpx * secG) >> 1;
    int tmpy = sAbsGY == 0 ? 0 : shift_msb(sDIY * 4 - tmpData, sAbsGY);
    tmpy = clip3(-limit, limit, tmpy);
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {          /* addBIOAvg4 :109 */
      const int i = (by + y + 1) * S + bx + x + 1;
      const int b = tmpx * (gx[0][i] - gx[1][i]) + tmpy * (gy[0][i] - gy[1][i]);
      dst[(by + y) * ds + bx + x] = (int16_t)clip3(0, pmax, (int16_t)((Q[0][i] + Q[1][i] + b + offset) >> shiftNum));
    }
  }
  for (int l = 0; l < 2; l++) { free(gx[l]); free(gy[l]); free(Q[l]); }
}

/* ---- DMVR (:1847 xProcessDMVR) ---- */
/* bilinear prediction for the search (xinitMC :1804 -> xPredInterBlk bilinearMC -> filter<2> :556 / filterCopy biMCForDMVR :445), 10-bit */
static void bilinear(const b200_geom* g, const Plane* ref, int x0, int y0, int xF, int yF, int w, int h, int16_t* out, int os)
{
  const int bd = g->bitDepth, sh1 = 4 - (10 - bd), o1 = 1 << (sh1 - 1);
  const int8_t* fh = &kIfBilin4[xF * 2]; const int8_t* fv = &kIfBilin4[yF * 2];
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    int v;
    if (xF == 0 && yF == 0) v = px(ref, x0 + x, y0 + y) * (1 << (10 - bd));
    else if (yF == 0) v = (fh[0] * px(ref, x0 + x, y0 + y) + fh[1] * px(ref, x0 + x + 1, y0 + y) + o1) >> sh1;
    else if (xF == 0) v = (fv[0] * px(ref, x0 + x, y0 + y) + fv[1] * px(ref, x0 + x, y0 + y + 1) + o1) >> sh1;
    else {
      const int a = (int16_t)((fh[0] * px(ref, x0 + x, y0 + y) + fh[1] * px(ref, x0 + x + 1, y0 + y) + o1) >> sh1);
      const int b = (int16_t)((fh[0] * px(ref, x0 + x, y0 + y + 1) + fh[1] * px(ref, x0 + x + 1, y0 + y + 1) + o1) >> sh1);
      v = (fv[0] * a + fv[1] * b + 8) >> 4;
    }
    out[y * os + x] = (int16_t)v;
  }
}

static unsigned sad_even_rows(const int16_t* a, const int16_t* b, int stride, int w, int h)
{ unsigned s = 0; for (int y = 0; y < h; y += 2) for (int x = 0; x < w; x++) s += iabs(a[y * stride + x] - b[y * stride + x]); return s; }

static int div_for_maxq7(int64_t N, int64_t D)     /* :1612 */
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

/* final MC of one list for a DMVR sub-block (xFinalPaddedMCForDMVR :1731 + xPrefetchPad :1525 + prefetchPadCore :311).
 * When the refined MV changes the integer position the reference window of the ORIGINAL merge MV is used, padded by 2 (1 for
 * 4:2:0 chroma) replicated samples: reads outside that window see replicas, not the real picture. */
typedef struct { Plane base; int wx0, wy0, ww, wh; } Win;   /* window = rectangle of `base` whose outside is replicated */
static inline int pxw(const Win* W, int x, int y) { return px(&W->base, clip3(W->wx0, W->wx0 + W->ww - 1, x), clip3(W->wy0, W->wy0 + W->wh - 1, y)); }

static void interp_win(const Win* W, int x0, int y0, const int8_t* fh, const int8_t* fv, int taps, int w, int h, int bd, int16_t* out, int os)
{
  const int half = taps / 2 - 1, headroom = 14 - bd < 2 ? 2 : 14 - bd, sh1 = 6 - headroom;
  for (int x = 0; x < w; x++) {
    int col[16 + 8];
    for (int y = 0; y < h + taps - 1; y++) {
      int s = 0;
      for (int t = 0; t < taps; t++) s += fh[t] * pxw(W, x0 + x + t - half, y0 + y - half);
      col[y] = (int16_t)((s - (IF_OFFS << sh1)) >> sh1);
    }
    for (int y = 0; y < h; y++) { int s = 0; for (int t = 0; t < taps; t++) s += fv[t] * col[y + t]; out[y * os + x] = (int16_t)(s >> 6); }
  }
}

static void dmvr_final_list(const b200_geom* g, const Plane ref[3], int sx, int sy, int dx, int dy, const int mvRef[2] /*refined*/, const int mvMerge[2],
                            int altHpel, int bio, int16_t* outY /* bio: (dx+2)x(dy+2) ring buffer, else dx x dy */, int16_t* outC[2])
{
  int mvc[2] = { mvRef[0], mvRef[1] };
  clip_mv(mvc, sx, sy, g);                                               /* cMvClipped, relative to the sub-block (:1749) */
  for (int comp = 0; comp < (g->chromaFormat ? 3 : 1); comp++) {
    const int sh = comp ? 5 : 4, taps = comp ? 4 : 8, cs = comp ? 1 : 0;
    const int dIx = (mvRef[0] >> sh) - (mvMerge[0] >> sh), dIy = (mvRef[1] >> sh) - (mvMerge[1] >> sh);
    const int bw = dx >> cs, bh = dy >> cs, bx = sx >> cs, by = sy >> cs;
    const int xF = mvc[0] & ((1 << sh) - 1), yF = mvc[1] & ((1 << sh) - 1);
    const int8_t *fh, *fv;
    if (comp == 0) { fh = luma_taps(xF, bw, bh, altHpel); fv = luma_taps(yF, bw, bh, altHpel); }
    else { fh = &kIfChroma[xF * 4]; fv = &kIfChroma[yF * 4]; }
    int16_t* out = comp == 0 ? (bio ? outY + (dx + 2) + 1 : outY) : outC[comp - 1];
    const int os = comp == 0 ? (bio ? dx + 2 : dx) : bw;
    Win W; W.base = ref[comp];
    int ox, oy;                                                          /* integer position of output sample (0,0) in the reference */
    if (dIx || dIy) {
      /* xPrefetchPad: window fetched for (merge MV - (taps/2-1)) clipped relative to the sub-block, size (bw+taps-1)x(bh+taps-1) */
      int pm[2] = { mvMerge[0] - ((taps / 2 - 1) << sh), mvMerge[1] - ((taps / 2 - 1) << sh) };
      clip_mv(pm, sx, sy, g);
      W.wx0 = bx + (pm[0] >> sh); W.wy0 = by + (pm[1] >> sh); W.ww = bw + taps - 1; W.wh = bh + taps - 1;
      ox = W.wx0 + (taps / 2 - 1) + dIx; oy = W.wy0 + (taps / 2 - 1) + dIy;
    } else {
      W.wx0 = -100000; W.wy0 = -100000; W.ww = 400000; W.wh = 400000;   /* no window: plain picture access */
      ox = bx + (mvc[0] >> sh); oy = by + (mvc[1] >> sh);
    }
    interp_win(&W, ox, oy, fh, fv, taps, bw, bh, g->bitDepth, out, os);
    if (comp == 0 && bio) {                                              /* BDOF ring from the same (possibly padded) source, :847-885 */
      const int S = dx + 2, shift = 14 - g->bitDepth < 2 ? 2 : 14 - g->bitDepth;
      const int rx = ox - (xF < 8 ? 1 : 0), ry = oy - (yF < 8 ? 1 : 0);
      for (int y = 0; y < dy + 2; y++) for (int x = 0; x < dx + 2; x++) {
        if (x > 0 && x < dx + 1 && y > 0 && y < dy + 1) continue;
        outY[y * S + x] = (int16_t)((int16_t)(pxw(&W, rx + x, ry + y) << shift) - IF_OFFS);
      }
    }
  }
}

static void dmvr_pu(const b200_geom* g, const b200_pu* pu, const Plane r0[3], const Plane r1[3], int16_t* const dst[3], int32_t* dmvrMv)
{
  const int w = pu->w, h = pu->h, bd = g->bitDepth, altHpel = (pu->flags & B200_PU_ALTHPEL) != 0;
  const int S = w + 4;
  int16_t* L0 = (int16_t*)malloc((size_t)S * (h + 4) * 2); int16_t* L1 = (int16_t*)malloc((size_t)S * (h + 4) * 2);
  const int mrg[2][2] = { { pu->mv[0][0], pu->mv[0][1] }, { pu->mv[1][0], pu->mv[1][1] } };
  for (int l = 0; l < 2; l++) {                                           /* xinitMC */
    int m[2] = { mrg[l][0], mrg[l][1] };
    clip_mv(m, pu->x, pu->y, g);
    m[0] -= 2 << 4; m[1] -= 2 << 4;
    bilinear(g, l ? &r1[0] : &r0[0], pu->x + (m[0] >> 4), pu->y + (m[1] >> 4), m[0] & 15, m[1] & 15, w + 4, h + 4, l ? L1 : L0, S);
  }
  const int dx = w < 16 ? w : 16, dy = h < 16 ? h : 16;
  int num = 0;
  int16_t P0[18 * 18], P1[18 * 18], c0[2][64], c1[2][64];
  for (int ys = 0; ys < h; ys += dy) for (int xs = 0; xs < w; xs += dx, num++) {
    const int16_t* b0 = L0 + (2 + ys) * S + 2 + xs; const int16_t* b1 = L1 + (2 + ys) * S + 2 + xs;
    unsigned minCost = sad_even_rows(b0, b1, S, dx, dy);                  /* distFunc = SAD<<1, then >>=1 (:1921-1924) */
    minCost -= minCost >> 2;
    int dmv[2] = { 0, 0 };
    if (minCost >= (unsigned)(dx * dy)) {
      unsigned sads[25]; sads[12] = minCost;
      int best[2] = { 0, 0 };
      for (int v = -2; v <= 2; v++) for (int u = -2; u <= 2; u++) {        /* xBIPMVRefine :1702 */
        if (!(u == 0 && v == 0)) sads[(v + 2) * 5 + u + 2] = sad_even_rows(b0 + v * S + u, b1 - v * S - u, S, dx, dy);
        if (sads[(v + 2) * 5 + u + 2] < minCost) { minCost = sads[(v + 2) * 5 + u + 2]; best[0] = u; best[1] = v; }
      }
      dmv[0] = best[0] * 16; dmv[1] = best[1] * 16;
      if (iabs(dmv[0]) != 32 && iabs(dmv[1]) != 32) {                     /* xDMVRSubPixelErrorSurface :1785 / xSubPelErrorSrfc :1647 */
        const unsigned* c = &sads[(best[1] + 2) * 5 + best[0] + 2];
        const uint64_t sb[5] = { c[0], c[-1], c[-5], c[1], c[5] };
        for (int d = 0; d < 2; d++) {
          const uint64_t a = sb[1 + d], b = sb[3 + d];
          const int64_t num64 = (int64_t)(a - b) * 16, den = (int64_t)(a + b - (sb[0] << 1));
          if (den != 0) {
            if (a != sb[0] && b != sb[0]) dmv[d] += div_for_maxq7(num64, den);
            else dmv[d] += (a == sb[0]) ? -8 : 8;
          }
        }
      }
    }
    if (dmvrMv) { dmvrMv[(pu->dmvrOff + num) * 2] = dmv[0]; dmvrMv[(pu->dmvrOff + num) * 2 + 1] = dmv[1]; }
    int mv0[2] = { clip3(-(1 << 17), (1 << 17) - 1, mrg[0][0] + dmv[0]), clip3(-(1 << 17), (1 << 17) - 1, mrg[0][1] + dmv[1]) };
    int mv1[2] = { clip3(-(1 << 17), (1 << 17) - 1, mrg[1][0] - dmv[0]), clip3(-(1 << 17), (1 << 17) - 1, mrg[1][1] - dmv[1]) };
    const int early = minCost < (unsigned)(dx * dy) && dmv[0] == 0 && dmv[1] == 0;
    (void)early;
    const int bio = (pu->flags & B200_PU_BDOF) && !(minCost < (unsigned)(2 * dx * dy));
    int16_t* cc0[2] = { c0[0], c0[1] }; int16_t* cc1[2] = { c1[0], c1[1] };
    dmvr_final_list(g, r0, pu->x + xs, pu->y + ys, dx, dy, mv0, mrg[0], altHpel, bio, P0, cc0);
    dmvr_final_list(g, r1, pu->x + xs, pu->y + ys, dx, dy, mv1, mrg[1], altHpel, bio, P1, cc1);
    int16_t* dY = dst[0] + (size_t)(pu->y + ys) * g->stride[0] + pu->x + xs;
    if (bio) bdof(P0, P1, dx, dy, bd, dY, g->stride[0]); else average(P0, P1, dx, dy, dx, bd, 4, dY, g->stride[0]);
    if (g->chromaFormat) for (int c = 0; c < 2; c++)
      average(c0[c], c1[c], dx >> 1, dy >> 1, dx >> 1, bd, 4, dst[1 + c] + (size_t)((pu->y + ys) >> 1) * g->stride[1 + c] + ((pu->x + xs) >> 1), g->stride[1 + c]);
  }
  free(L0); free(L1);
}

/* ---- affine (:934 xPredAffineBlk) ---- */
static void round_affine(int* x, int* y, int s) { const int o = 1 << (s - 1); *x = (*x + o - (*x >= 0)) >> s; *y = (*y + o - (*y >= 0)) >> s; }   /* Mv.cpp:57 */

static int spread_over_limit(int a, int b, int c, int d, int predType)    /* :892 */
{
  const int s4 = 4 << 11, ft = 6;
  #define MX(p,q) ((p) > (q) ? (p) : (q))
  #define MN(p,q) ((p) < (q) ? (p) : (q))
  if (predType == 3) {
    int rw = MX(MX(0, 4 * a + s4), MX(4 * c, 4 * a + 4 * c + s4)) - MN(MN(0, 4 * a + s4), MN(4 * c, 4 * a + 4 * c + s4));
    int rh = MX(MX(0, 4 * b), MX(4 * d + s4, 4 * b + 4 * d + s4)) - MN(MN(0, 4 * b), MN(4 * d + s4, 4 * b + 4 * d + s4));
    rw = (rw >> 11) + ft + 3; rh = (rh >> 11) + ft + 3;
    return rw * rh > (ft + 9) * (ft + 9);
  }
  int rw = MX(0, 4 * a + s4) - MN(0, 4 * a + s4), rh = MX(0, 4 * b) - MN(0, 4 * b);
  rw = (rw >> 11) + ft + 3; rh = (rh >> 11) + ft + 3;
  if (rw * rh > (ft + 9) * (ft + 5)) return 1;
  rw = MX(0, 4 * c) - MN(0, 4 * c); rh = MX(0, 4 * d + s4) - MN(0, 4 * d + s4);
  rw = (rw >> 11) + ft + 3; rh = (rh >> 11) + ft + 3;
  return rw * rh > (ft + 5) * (ft + 9);
  #undef MX
  #undef MN
}

/* one list of an affine PU into 14-bit (bi) or final (uni) buffers for all components */
static void affine_list(const b200_geom* g, const b200_pu* pu, int l, const Plane ref[3], int bi, int16_t* out[3], const int os[3])
{
  const int w = pu->w, h = pu->h, bd = g->bitDepth, pmax = (1 << bd) - 1;
  int l2w = 0, l2h = 0; while ((1 << l2w) < w) l2w++; while ((1 << l2h) < h) l2h++;
  const int LT[2] = { pu->mv[l][0], pu->mv[l][1] }, RT[2] = { pu->cpmv[l][0][0], pu->cpmv[l][0][1] }, LB[2] = { pu->cpmv[l][1][0], pu->cpmv[l][1][1] };
  const int six = (pu->flags & B200_PU_AFFINE6) != 0;
  const int dHX = (RT[0] - LT[0]) * (1 << (7 - l2w)), dHY = (RT[1] - LT[1]) * (1 << (7 - l2w));
  const int dVX = six ? (LB[0] - LT[0]) * (1 << (7 - l2h)) : -dHY, dVY = six ? (LB[1] - LT[1]) * (1 << (7 - l2h)) : dHX;
  const int over = spread_over_limit(dHX, dHY, dVX, dVY, pu->interDir);
  int prof = (pu->flags & (l ? B200_PU_PROF1 : B200_PU_PROF0)) != 0;
  if (six ? (LT[0] == RT[0] && LT[1] == RT[1] && LT[0] == LB[0] && LT[1] == LB[1]) : (LT[0] == RT[0] && LT[1] == RT[1])) prof = 0;
  if (over) prof = 0;
  const int last = prof ? 0 : !bi;
  int dMvH[16], dMvV[16];
  if (prof) {
    const int qHX = dHX * 4, qHY = dHY * 4, qVX = dVX * 4, qVY = dVY * 4;
    dMvH[0] = ((dHX + dVX) * 2) - ((qHX + qVX) * 2); dMvV[0] = ((dHY + dVY) * 2) - ((qHY + qVY) * 2);
    for (int i = 1; i < 4; i++) { dMvH[i] = dMvH[i - 1] + qHX; dMvV[i] = dMvV[i - 1] + qHY; }
    for (int j = 1; j < 4; j++) for (int i = 0; i < 4; i++) { dMvH[j * 4 + i] = dMvH[(j - 1) * 4 + i] + qVX; dMvV[j * 4 + i] = dMvV[(j - 1) * 4 + i] + qVY; }
    for (int i = 0; i < 16; i++) { round_affine(&dMvH[i], &dMvV[i], 8); dMvH[i] = clip3(-31, 31, dMvH[i]); dMvV[i] = clip3(-31, 31, dMvV[i]); }
  }
  const int hMax = (g->width + 8 - pu->x - 1) * 16, hMin = (-g->ctuSize - 8 - pu->x + 1) * 16;
  const int vMax = (g->height + 8 - pu->y - 1) * 16, vMin = (-g->ctuSize - 8 - pu->y + 1) * 16;
  const int nbx = w / 4, nby = h / 4;
  int* mvf = (int*)malloc(sizeof(int) * 2 * nbx * nby);
  for (int j = 0; j < nby; j++) for (int i = 0; i < nbx; i++) {           /* PU::setAllAffineMv UnitTools.cpp:2689 */
    int mx, my;
    if (over) { mx = LT[0] * 128 + dHX * (w >> 1) + dVX * (h >> 1); my = LT[1] * 128 + dHY * (w >> 1) + dVY * (h >> 1); }
    else      { mx = LT[0] * 128 + dHX * (2 + 4 * i) + dVX * (2 + 4 * j); my = LT[1] * 128 + dHY * (2 + 4 * i) + dVY * (2 + 4 * j); }
    round_affine(&mx, &my, 7);
    mvf[(j * nbx + i) * 2] = clip3(-(1 << 17), (1 << 17) - 1, mx); mvf[(j * nbx + i) * 2 + 1] = clip3(-(1 << 17), (1 << 17) - 1, my);
  }
  const int shift = 14 - bd < 2 ? 2 : 14 - bd;
  for (int j = 0; j < nby; j++) for (int i = 0; i < nbx; i++) {           /* luma 4x4 sub-blocks, 6-tap filters */
    int mv[2] = { clip3(hMin, hMax, mvf[(j * nbx + i) * 2]), clip3(vMin, vMax, mvf[(j * nbx + i) * 2 + 1]) };
    const int bx = pu->x + 4 * i, by = pu->y + 4 * j;
    int16_t* o = out[0] + (4 * j) * os[0] + 4 * i;
    if (!prof) { pred_block(g, &ref[0], 0, bx, by, 4, 4, mv, 0, last, o, os[0]); continue; }
    int16_t E[36];                                                        /* 6x6: prediction + ring of integer samples (:1233-1262) */
    pred_block(g, &ref[0], 0, bx, by, 4, 4, mv, 0, 0, E + 7, 6);
    const int xF = mv[0] & 15, yF = mv[1] & 15;
    const int rx = bx + (mv[0] >> 4) + (xF >> 3) - 1, ry = by + (mv[1] >> 4) + (yF >> 3) - 1;
    for (int y = 0; y < 6; y++) for (int x = 0; x < 6; x++) {
      if (x > 0 && x < 5 && y > 0 && y < 5) continue;
      E[y * 6 + x] = (int16_t)((int16_t)(px(&ref[0], rx + x, ry + y) << shift) - IF_OFFS);
    }
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {             /* gradFilterCore<false> :212 + applyPROFCore :61 */
      const int c = (y + 1) * 6 + x + 1;
      const int gX = (E[c + 1] >> 6) - (E[c - 1] >> 6), gY = (E[c + 6] >> 6) - (E[c - 6] >> 6);
      const int lim = 1 << (bd + 1 > 13 ? bd + 1 : 13);
      int dI = clip3(-lim, lim - 1, dMvH[y * 4 + x] * gX + dMvV[y * 4 + x] * gY);
      int v = (int16_t)(E[c] + dI);
      if (!bi) v = clip3(0, pmax, (int16_t)((v + (1 << (shift - 1)) + IF_OFFS) >> shift));
      o[y * os[0] + x] = (int16_t)v;
    }
  }
  if (g->chromaFormat) for (int j = 0; j < nby / 2; j++) for (int i = 0; i < nbx / 2; i++) {   /* chroma 4x4 = luma 8x8: MV = avg of TL and BR luma sub-blocks (:1135-1151) */
    const int* a = &mvf[((2 * j) * nbx + 2 * i) * 2]; const int* b = &mvf[((2 * j + 1) * nbx + 2 * i + 1) * 2];
    int mx = (a[0] + b[0]), my = (a[1] + b[1]);
    round_affine(&mx, &my, 1);
    int mv[2] = { clip3(hMin, hMax, mx), clip3(vMin, vMax, my) };
    for (int c = 1; c < 3; c++) pred_block(g, &ref[c], c, pu->x + 8 * i, pu->y + 8 * j, 8, 8, mv, 0, !bi, out[c] + (4 * j) * os[c] + 4 * i, os[c]);
  }
  free(mvf);
}

/* ---- PU driver (motionCompensation :1372) ---- */
void orc_mc_predict(const b200_geom* g, int16_t* const dst[3], const int16_t* const* refs, const b200_pu* pus, size_t numPus, int32_t* dmvrMv)
{
  orc_mc_predict_wp(g, dst, refs, pus, numPus, dmvrMv, NULL);
}

void orc_mc_predict_wp(const b200_geom* g, int16_t* const dst[3], const int16_t* const* refs, const b200_pu* pus, size_t numPus, int32_t* dmvrMv, const b200_wp* wp)
{
  const int nComp = g->chromaFormat ? 3 : 1, bd = g->bitDepth;
  int16_t* tmp[2][3];
  for (int l = 0; l < 2; l++) for (int c = 0; c < 3; c++) tmp[l][c] = (int16_t*)malloc(130 * 130 * 2);
  for (size_t n = 0; n < numPus; n++) {
    const b200_pu* pu = &pus[n];
    Plane R[2][3];
    for (int l = 0; l < 2; l++) for (int c = 0; c < nComp; c++) if (pu->refSlot[l] >= 0) {
      R[l][c].p = refs[pu->refSlot[l] * 3 + c]; R[l][c].w = c ? g->width >> 1 : g->width; R[l][c].h = c ? g->height >> 1 : g->height; R[l][c].stride = g->stride[c];
    }
    const int bi = pu->refSlot[0] >= 0 && pu->refSlot[1] >= 0, altHpel = (pu->flags & B200_PU_ALTHPEL) != 0;
    int16_t* d[3]; for (int c = 0; c < nComp; c++) d[c] = dst[c] + (size_t)(pu->y >> (c ? 1 : 0)) * g->stride[c] + (pu->x >> (c ? 1 : 0));
    if (pu->flags & B200_PU_DMVR) { dmvr_pu(g, pu, R[0], R[1], dst, dmvrMv); continue; }
    if (pu->flags & B200_PU_GEO) {                                       /* motionCompensationGeo :1461: two uni-predictions of the whole CU, blended */
      int mvg[2][2];
      for (int l = 0; l < 2; l++) { mvg[l][0] = pu->mv[l][0]; mvg[l][1] = pu->mv[l][1]; clip_mv(mvg[l], pu->x, pu->y, g); }
      for (int c = 0; c < nComp; c++) {
        const int sw = pu->w >> (c ? 1 : 0), sh = pu->h >> (c ? 1 : 0);
        pred_block(g, &R[0][c], c, pu->x, pu->y, pu->w, pu->h, mvg[0], altHpel, 0, tmp[0][c], sw);
        pred_block(g, &R[1][c], c, pu->x, pu->y, pu->w, pu->h, mvg[1], altHpel, 0, tmp[1][c], sw);
        geo_blend(tmp[0][c], tmp[1][c], sw, sh, sw, bd, c, pu->w, pu->h, pu->bcwW1, d[c], g->stride[c]);
      }
      continue;
    }
    if (pu->flags & B200_PU_AFFINE) {
      const b200_wp* we = (wp && pu->wpIdx) ? &wp[pu->wpIdx - 1] : NULL;
      const int os[3] = { pu->w, pu->w >> 1, pu->w >> 1 };
      if (!bi) {
        const int l = pu->refSlot[0] >= 0 ? 0 : 1;
        if (!we) { affine_list(g, pu, l, R[l], 0, d, g->stride); continue; }
        affine_list(g, pu, l, R[l], 1, tmp[0], os);                      /* xPredInterUni(bi = true) + xWeightedPredictionBi/Uni (InterPrediction.cpp:707-741) */
        for (int c = 0; c < nComp; c++) weighted(tmp[0][c], NULL, pu->w >> (c ? 1 : 0), pu->h >> (c ? 1 : 0), os[c], bd, we, c, d[c], g->stride[c]);
        continue;
      }
      affine_list(g, pu, 0, R[0], 1, tmp[0], os); affine_list(g, pu, 1, R[1], 1, tmp[1], os);
      for (int c = 0; c < nComp; c++) {
        if (we) weighted(tmp[0][c], tmp[1][c], pu->w >> (c ? 1 : 0), pu->h >> (c ? 1 : 0), os[c], bd, we, c, d[c], g->stride[c]);
        else    average(tmp[0][c], tmp[1][c], pu->w >> (c ? 1 : 0), pu->h >> (c ? 1 : 0), os[c], bd, pu->bcwW1, d[c], g->stride[c]);
      }
      continue;
    }
    int mv[2][2];
    for (int l = 0; l < 2; l++) { mv[l][0] = pu->mv[l][0]; mv[l][1] = pu->mv[l][1]; clip_mv(mv[l], pu->x, pu->y, g); }
    const b200_wp* we = (wp && pu->wpIdx) ? &wp[pu->wpIdx - 1] : NULL;
    if (!bi) {
      const int l = pu->refSlot[0] >= 0 ? 0 : 1;
      for (int c = 0; c < nComp; c++) {
        if (!we) { pred_block(g, &R[l][c], c, pu->x, pu->y, pu->w, pu->h, mv[l], altHpel, 1, d[c], g->stride[c]); continue; }
        const int sw = pu->w >> (c ? 1 : 0), sh = pu->h >> (c ? 1 : 0);
        pred_block(g, &R[l][c], c, pu->x, pu->y, pu->w, pu->h, mv[l], altHpel, 0, tmp[0][c], sw);
        weighted(tmp[0][c], NULL, sw, sh, sw, bd, we, c, d[c], g->stride[c]);
      }
      continue;
    }
    if (pu->flags & B200_PU_BDOF) {                                       /* xSubPuBio :551: <=16x16 sub-blocks, MV clipped relative to the CU */
      const int sw = pu->w < 16 ? pu->w : 16, sh = pu->h < 16 ? pu->h : 16;
      int16_t P0[18 * 18], P1[18 * 18];
      for (int y = 0; y < pu->h; y += sh) for (int x = 0; x < pu->w; x += sw) {
        bdof_fill(g, &R[0][0], pu->x + x, pu->y + y, sw, sh, mv[0], altHpel, P0);
        bdof_fill(g, &R[1][0], pu->x + x, pu->y + y, sw, sh, mv[1], altHpel, P1);
        bdof(P0, P1, sw, sh, bd, d[0] + y * g->stride[0] + x, g->stride[0]);
      }
      for (int c = 1; c < nComp; c++) {
        pred_block(g, &R[0][c], c, pu->x, pu->y, pu->w, pu->h, mv[0], altHpel, 0, tmp[0][c], pu->w >> 1);
        pred_block(g, &R[1][c], c, pu->x, pu->y, pu->w, pu->h, mv[1], altHpel, 0, tmp[1][c], pu->w >> 1);
        average(tmp[0][c], tmp[1][c], pu->w >> 1, pu->h >> 1, pu->w >> 1, bd, 4, d[c], g->stride[c]);
      }
      continue;
    }
    for (int c = 0; c < nComp; c++) {                                     /* xPredInterBi :686 */
      const int sw = pu->w >> (c ? 1 : 0), sh = pu->h >> (c ? 1 : 0);
      pred_block(g, &R[0][c], c, pu->x, pu->y, pu->w, pu->h, mv[0], altHpel, 0, tmp[0][c], sw);
      pred_block(g, &R[1][c], c, pu->x, pu->y, pu->w, pu->h, mv[1], altHpel, 0, tmp[1][c], sw);
      if (we) weighted(tmp[0][c], tmp[1][c], sw, sh, sw, bd, we, c, d[c], g->stride[c]);
      else    average(tmp[0][c], tmp[1][c], sw, sh, sw, bd, pu->bcwW1, d[c], g->stride[c]);
    }
  }
  for (int l = 0; l < 2; l++) for (int c = 0; c < 3; c++) free(tmp[l][c]);
}
