/* oracle/k5_alf.c — CPU restatement of K5 (ALF classification, 7x7 / 5x5 diamond filters, CC-ALF).
 * TEST INFRASTRUCTURE ONLY — see vvc_oracle.h. Pinned against oracle/_ref (tests/test_k45_oracle_vs_ref.py). */
#include "vvc_oracle.h"
#include <stdlib.h>
#include <string.h>

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* AdaptiveLoopFilter.cpp:969-1173.  Laplacians on the 2x2-subsampled grid over the 8x8 window of each 4x4 block. */
void orc_alf_classify(uint16_t* cls, const int16_t* src, ptrdiff_t stride, int blkX, int blkY, int blkW, int blkH,
                      int shift, int vbH, int vbPos)
{
  static const int th[16] = { 0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4 };
  static const int transposeTable[8] = { 0, 1, 0, 2, 2, 3, 1, 3 };
  for (int by = 0; by < blkH; by += 4) for (int bx = 0; bx < blkW; bx += 4) {
    const int y0 = blkY + by, x0 = blkX + bx;
    int sum[4] = { 0, 0, 0, 0 };          /* V, H, D0, D1 */
    const int aboveVb = (y0 % vbH) == vbPos - 4, belowVb = (y0 % vbH) == vbPos;
    for (int r = 0; r < 4; r++) {         /* row pairs: laplacian rows y = y0-2+2r (and the +1 partner) */
      if ((aboveVb && r == 3) || (belowVb && r == 0)) continue;
      const int y = y0 - 2 + 2 * r;
      ptrdiff_t up = -stride, dn2 = 2 * stride;                 /* rows y-1 and y+2 */
      if (y > 0 && (y % vbH) == vbPos - 2) dn2 = stride;        /* :1006-1015 */
      else if (y > 0 && (y % vbH) == vbPos) up = 0;
      for (int c = 0; c < 4; c++) {
        const int x = x0 - 2 + 2 * c;
        const int16_t* p = src + (ptrdiff_t)y * stride + x;     /* p = (x,y); partner sample = (x+1,y+1) */
        const int a = p[0] << 1, b = p[stride + 1] << 1;
        sum[0] += iabs(a - p[up] - p[stride]) + iabs(b - p[1] - p[dn2 + 1]);
        sum[1] += iabs(a - p[1] - p[-1]) + iabs(b - p[stride + 2] - p[stride]);
        sum[2] += iabs(a - p[up - 1] - p[stride + 1]) + iabs(b - p[0] - p[dn2 + 2]);
        sum[3] += iabs(a - p[stride - 1] - p[up + 1]) + iabs(b - p[dn2] - p[2]);
      }
    }
    const int sumV = sum[0], sumH = sum[1], sumD0 = sum[2], sumD1 = sum[3];
    const int act = clip3(0, 15, ((sumV + sumH) * ((aboveVb || belowVb) ? 96 : 64)) >> shift);
    int classIdx = th[act];
    int hv1, hv0, d1, d0, dirHV, dirD;
    if (sumV > sumH) { hv1 = sumV; hv0 = sumH; dirHV = 1; } else { hv1 = sumH; hv0 = sumV; dirHV = 3; }
    if (sumD0 > sumD1) { d1 = sumD0; d0 = sumD1; dirD = 0; } else { d1 = sumD1; d0 = sumD0; dirD = 2; }
    int hvd1, hvd0, mainDir, secDir;
    if ((uint32_t)d1 * (uint32_t)hv0 > (uint32_t)hv1 * (uint32_t)d0) { hvd1 = d1; hvd0 = d0; mainDir = dirD; secDir = dirHV; }
    else { hvd1 = hv1; hvd0 = hv0; mainDir = dirHV; secDir = dirD; }
    int strength = 0;
    if (hvd1 > 2 * hvd0) strength = 1;
    if (hvd1 * 2 > 9 * hvd0) strength = 2;
    if (strength) classIdx += (((mainDir & 1) << 1) + strength) * 5;
    cls[(by / 4) * 8 + bx / 4] = (uint16_t)(classIdx | (transposeTable[mainDir * 2 + (secDir >> 1)] << 8));
  }
}

static inline int clip_alf(int clip, int ref, int v0, int v1) { return clip3(-clip, clip, v0 - ref) + clip3(-clip, clip, v1 - ref); }

/* AdaptiveLoopFilter.cpp:1175-1346 */
void orc_alf_filter_blk(int is7, const uint16_t* cls, int16_t* dst, ptrdiff_t ds, const int16_t* src, ptrdiff_t ss,
                        int blkX, int blkY, int blkW, int blkH, const int16_t* coeffSet, const int16_t* clipSet, int bd, int vbH, int vbPos)
{
  const int pmax = (1 << bd) - 1, reach = is7 ? 3 : 2;
  for (int y = blkY; y < blkY + blkH; y++) {
    const int yVb = y & (vbH - 1);
    /* rows are clamped symmetrically so that no tap crosses the virtual boundary (:1254-1273) */
    int lim = reach;
    if (yVb < vbPos && yVb >= vbPos - (is7 ? 4 : 2)) lim = vbPos - 1 - yVb;
    else if (yVb >= vbPos && yVb <= vbPos + (is7 ? 3 : 1)) lim = yVb - vbPos;
    const int nearVb = yVb == vbPos - 1 || yVb == vbPos;
    for (int x = blkX; x < blkX + blkW; x++) {
      const int16_t* f = coeffSet; const int16_t* c = clipSet;
      if (cls) {
        const uint16_t k = cls[((y - blkY) / 4) * 8 + (x - blkX) / 4];
        const int off = (k & 0xff) * 13 + (k >> 8) * 13 * 25;
        f += off; c += off;
      }
      const int16_t* p = src + (ptrdiff_t)y * ss + x;
      #define ROW(k) ((ptrdiff_t)((k) < lim ? (k) : lim) * ss)
      const int cur = p[0];
      int sum = 0;
      if (is7) {
        sum += f[0]  * clip_alf(c[0],  cur, p[ROW(3)],      p[-ROW(3)]);
        sum += f[1]  * clip_alf(c[1],  cur, p[ROW(2) + 1],  p[-ROW(2) - 1]);
        sum += f[2]  * clip_alf(c[2],  cur, p[ROW(2)],      p[-ROW(2)]);
        sum += f[3]  * clip_alf(c[3],  cur, p[ROW(2) - 1],  p[-ROW(2) + 1]);
        sum += f[4]  * clip_alf(c[4],  cur, p[ROW(1) + 2],  p[-ROW(1) - 2]);
        sum += f[5]  * clip_alf(c[5],  cur, p[ROW(1) + 1],  p[-ROW(1) - 1]);
        sum += f[6]  * clip_alf(c[6],  cur, p[ROW(1)],      p[-ROW(1)]);
        sum += f[7]  * clip_alf(c[7],  cur, p[ROW(1) - 1],  p[-ROW(1) + 1]);
        sum += f[8]  * clip_alf(c[8],  cur, p[ROW(1) - 2],  p[-ROW(1) + 2]);
        sum += f[9]  * clip_alf(c[9],  cur, p[3],  p[-3]);
        sum += f[10] * clip_alf(c[10], cur, p[2],  p[-2]);
        sum += f[11] * clip_alf(c[11], cur, p[1],  p[-1]);
      } else {
        sum += f[0] * clip_alf(c[0], cur, p[ROW(2)],     p[-ROW(2)]);
        sum += f[1] * clip_alf(c[1], cur, p[ROW(1) + 1], p[-ROW(1) - 1]);
        sum += f[2] * clip_alf(c[2], cur, p[ROW(1)],     p[-ROW(1)]);
        sum += f[3] * clip_alf(c[3], cur, p[ROW(1) - 1], p[-ROW(1) + 1]);
        sum += f[4] * clip_alf(c[4], cur, p[2], p[-2]);
        sum += f[5] * clip_alf(c[5], cur, p[1], p[-1]);
      }
      #undef ROW
      sum = nearVb ? (sum + (1 << 9)) >> 10 : (sum + 64) >> 7;
      dst[(ptrdiff_t)y * ds + x] = (int16_t)clip3(0, pmax, sum + cur);
    }
  }
}

/* AdaptiveLoopFilter.cpp:1348-1445, 4:2:0 */
void orc_alf_ccalf_blk(int16_t* dstC, ptrdiff_t cs, const int16_t* srcL, ptrdiff_t ls, int cX, int cY, int cW, int cH,
                       const int16_t* f, int bd, int vbH, int vbPos)
{
  const int pmax = (1 << bd) - 1, half = (1 << bd) >> 1;
  for (int y = cY; y < cY + cH; y++) {
    const int pos = (y << 1) & (vbH - 1);
    ptrdiff_t o1 = ls, o2 = -ls, o3 = 2 * ls;
    if (pos == vbPos - 2 || pos == vbPos + 1) o3 = o1;
    else if (pos == vbPos - 1 || pos == vbPos) o1 = o2 = o3 = 0;
    for (int x = cX; x < cX + cW; x++) {
      const int16_t* l = srcL + (ptrdiff_t)(y << 1) * ls + (x << 1);
      const int cur = l[0];
      int sum = f[0] * (l[o2] - cur) + f[1] * (l[-1] - cur) + f[2] * (l[1] - cur) + f[3] * (l[o1 - 1] - cur)
              + f[4] * (l[o1] - cur) + f[5] * (l[o1 + 1] - cur) + f[6] * (l[o3] - cur);
      sum = (sum + 64) >> 7;
      sum = clip3(0, pmax, sum + half) - half;
      int16_t* d = dstC + (ptrdiff_t)y * cs + x;
      *d = (int16_t)clip3(0, pmax, sum + *d);
    }
  }
}

/* Picture level: AdaptiveLoopFilter.cpp:466 processCTU -> :664 filterCTU (no slice/tile/VB crossing). The source picture is
 * first copied into a buffer padded by 4 replicated samples, which is what prepareCTU (:453) does at picture borders. */
void orc_alf_picture(const b200_geom* g, const int16_t* const src[3], int16_t* const dst[3], const b200_alf_ctu* ctus,
                     const b200_alf_tables* T)
{
  const int PAD = 8;
  const int nComp = g->chromaFormat ? 3 : 1;
  int16_t* pad[3]; ptrdiff_t ps[3]; const int16_t* org[3];
  for (int c = 0; c < nComp; c++) {
    const int w = c ? g->width >> 1 : g->width, h = c ? g->height >> 1 : g->height;
    ps[c] = w + 2 * PAD;
    pad[c] = (int16_t*)malloc(sizeof(int16_t) * (size_t)ps[c] * (h + 2 * PAD));
    for (int y = -PAD; y < h + PAD; y++) for (int x = -PAD; x < w + PAD; x++) {
      const int sx = clip3(0, w - 1, x), sy = clip3(0, h - 1, y);
      pad[c][(size_t)(y + PAD) * ps[c] + x + PAD] = src[c][(size_t)sy * g->stride[c] + sx];
    }
    org[c] = pad[c] + (size_t)PAD * ps[c] + PAD;
  }
  const int ctu = g->ctuSize, ctusW = (g->width + ctu - 1) / ctu, ctusH = (g->height + ctu - 1) / ctu;
  const int vbH = ctu, vbPos = ctu - 4, vbHc = ctu >> 1, vbPosC = (ctu >> 1) - 2;
  uint16_t cls[64];
  int16_t* tmp[3] = { NULL, NULL, NULL };                      /* CTU + 8-sample frame, for CTUs with clipped sides / padded corners */
  const ptrdiff_t tstr = 128 + 16;
  for (int cy = 0; cy < ctusH; cy++) for (int cx = 0; cx < ctusW; cx++) {
    const b200_alf_ctu* p = &ctus[cy * ctusW + cx];
    const int x0 = cx * ctu, y0 = cy * ctu;
    const int w = x0 + ctu > g->width ? g->width - x0 : ctu, h = y0 + ctu > g->height ? g->height - y0 : ctu;
    const int16_t* srcC[3] = { org[0], org[1], org[2] }; ptrdiff_t strC[3] = { ps[0], ps[1], ps[2] };
    if (p->enable[0] & ~1) {
      /* filterCTU :763-848 with no virtual boundary inside the CTU: the CTU is copied with 4 (chroma 2) extra samples on the sides that may be read, the raster-slice
       * corners are padded, then the copy is extended by replication (extendBorderPel) — classification and filters run on the copy */
      const int f = p->enable[0];
      for (int c = 0; c < nComp; c++) {
        const int sh = c ? 1 : 0, bx = x0 >> sh, by = y0 >> sh, bw = w >> sh, bh = h >> sh, m = 4 >> sh;
        const int pl = (f & B200_ALF_CLIP_LEFT) || bx == 0 ? 0 : m, pr = (f & B200_ALF_CLIP_RIGHT) || bx + bw == (g->width >> sh) ? 0 : m;
        const int pt = (f & B200_ALF_CLIP_TOP) || by == 0 ? 0 : m, pb = (f & B200_ALF_CLIP_BOTTOM) || by + bh == (g->height >> sh) ? 0 : m;
        if (!tmp[c]) tmp[c] = (int16_t*)malloc(sizeof(int16_t) * (size_t)tstr * (128 + 16));
        int16_t* B = tmp[c] + 8 * tstr + 8;                      /* B[0] = sample (bx, by) */
        const int cw2 = bw + pl + pr, ch2 = bh + pt + pb;        /* the copied area starts at (-pl, -pt) */
        for (int y = -pt; y < bh + pb; y++) for (int x = -pl; x < bw + pr; x++) B[y * tstr + x] = org[c][(ptrdiff_t)(by + y) * ps[c] + bx + x];
        const int mg = (c && (p->enable[c] & B200_ALF_PAD_WIDE)) ? 4 : m;      /* padBorderPel margin: :794-803 pass the luma margin for a chroma plane on its own */
        int16_t* P0 = B - pt * tstr - pl;                        /* origin of the copied area */
        if (f & B200_ALF_PAD_TL) for (int y = 0; y < mg; y++) for (int x = 0; x < mg; x++) P0[y * tstr + x] = P0[y * tstr + mg];
        if (f & B200_ALF_PAD_BR) { int16_t* q = P0 + (ptrdiff_t)(ch2 - mg) * tstr + cw2 - mg; for (int y = 0; y < mg; y++) for (int x = 0; x < mg; x++) q[y * tstr + x] = q[y * tstr - 1]; }
        for (int y = 0; y < ch2; y++) for (int k = 1; k <= 4; k++) { P0[y * tstr - k] = P0[y * tstr]; P0[y * tstr + cw2 - 1 + k] = P0[y * tstr + cw2 - 1]; }     /* extendBorderPel(4) */
        for (int k = 1; k <= 4; k++) { memcpy(P0 - k * tstr - 4, P0 - 4, sizeof(int16_t) * (cw2 + 8)); memcpy(P0 + (ptrdiff_t)(ch2 - 1 + k) * tstr - 4, P0 + (ptrdiff_t)(ch2 - 1) * tstr - 4, sizeof(int16_t) * (cw2 + 8)); }
        srcC[c] = B - (ptrdiff_t)by * tstr - bx; strC[c] = tstr;  /* so that srcC[c][y * stride + x] is sample (x, y) of the picture */
      }
    }
    /* luma */
    if (p->enable[0] & 1) {
      const int16_t* coeff = T->lumaCoeff + (size_t)p->lumaSet * 4 * 25 * 13;
      const int16_t* clip  = T->lumaClip  + (size_t)p->lumaSet * 4 * 25 * 13;
      for (int by = 0; by < h; by += 32) for (int bx = 0; bx < w; bx += 32) {
        const int bw = bx + 32 > w ? w - bx : 32, bh = by + 32 > h ? h - by : 32;
        orc_alf_classify(cls, srcC[0], strC[0], x0 + bx, y0 + by, bw, bh, g->bitDepth + 4, vbH, vbPos);
        orc_alf_filter_blk(1, cls, dst[0], g->stride[0], srcC[0], strC[0], x0 + bx, y0 + by, bw, bh, coeff, clip, g->bitDepth, vbH, vbPos);
      }
    } else {
      for (int y = y0; y < y0 + h; y++) memcpy(dst[0] + (size_t)y * g->stride[0] + x0, src[0] + (size_t)y * g->stride[0] + x0, w * sizeof(int16_t));
    }
    /* chroma + CC-ALF */
    for (int c = 1; c < nComp; c++) {
      const int cx0 = x0 >> 1, cy0 = y0 >> 1, cw = w >> 1, chh = h >> 1;
      if (p->enable[c] & 1)
        orc_alf_filter_blk(0, NULL, dst[c], g->stride[c], srcC[c], strC[c], cx0, cy0, cw, chh, T->chromaCoeff + p->chromaAlt[c - 1] * 7,
                           T->chromaClip + p->chromaAlt[c - 1] * 7, g->bitDepth, vbHc, vbPosC);
      else
        for (int y = cy0; y < cy0 + chh; y++) memcpy(dst[c] + (size_t)y * g->stride[c] + cx0, src[c] + (size_t)y * g->stride[c] + cx0, cw * sizeof(int16_t));
      if (p->ccIdx[c - 1])
        orc_alf_ccalf_blk(dst[c], g->stride[c], srcC[0], strC[0], cx0, cy0, cw, chh, T->ccCoeff[c - 1] + (p->ccIdx[c - 1] - 1) * 7, g->bitDepth, vbH, vbPos);
    }
  }
  for (int c = 0; c < nComp; c++) { free(pad[c]); free(tmp[c]); }
}
