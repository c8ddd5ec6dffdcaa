/* oracle/k4_sao.c — CPU restatement of K4 (SAO). TEST INFRASTRUCTURE ONLY — see vvc_oracle.h.
 * Pinned against oracle/_ref (tests/test_k45_oracle_vs_ref.py). */
#include "vvc_oracle.h"
#include <string.h>

static inline int sgn(int v) { return (v > 0) - (v < 0); }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* SampleAdaptiveOffset.cpp:817 isProcessDisabled */
static int vb_disabled(int x, int y, int nV, const int* v, int nH, const int* h)
{
  for (int i = 0; i < nV; i++) if (x == v[i] || x == v[i] - 1) return 1;
  for (int i = 0; i < nH; i++) if (y == h[i] || y == h[i] - 1) return 1;
  return 0;
}

/* Which neighbouring region a sample (x,y) of a w x h block falls into -> availability bit (0 = inside the block). */
static unsigned region_bit(int x, int y, int w, int h)
{
  const int l = x < 0, r = x >= w, a = y < 0, b = y >= h;
  if (a) return l ? B200_AVAIL_AL : r ? B200_AVAIL_AR : B200_AVAIL_A;
  if (b) return l ? B200_AVAIL_BL : r ? B200_AVAIL_BR : B200_AVAIL_B;
  return l ? B200_AVAIL_L : r ? B200_AVAIL_R : 0;
}

/* SampleAdaptiveOffset.cpp:64-349. The reference walks lines with sign caches; per sample that is
 * edgeType = sgn(c - n0) + sgn(c - n1) with (n0,n1) the two neighbours along the EO direction, applied only when both
 * neighbours lie inside the block or in an available neighbouring CTU, and the sample is not next to a virtual boundary. */
void orc_sao_offset_block(int bitDepth, int typeIdx, const int* offset, const int16_t* src, int16_t* dst,
                          ptrdiff_t srcStride, ptrdiff_t dstStride, int width, int height, unsigned avail,
                          int numVerVb, const int* verVb, int numHorVb, const int* horVb)
{
  const int pmax = (1 << bitDepth) - 1;
  if (typeIdx == B200_SAO_BO) {
    const int shiftBits = bitDepth - 5;
    for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) {
      const int c = src[y * srcStride + x];
      dst[y * dstStride + x] = (int16_t)clip3(0, pmax, c + offset[c >> shiftBits]);
    }
    return;
  }
  static const int dx[4] = { 1, 0, 1, -1 }, dy[4] = { 0, 1, 1, 1 };   /* second neighbour; the first is the mirror */
  /* virtual boundaries: EO_0 checks vertical VBs only, EO_90 horizontal only, diagonals both (:106,:147,:196,:258) */
  const int nV = typeIdx == B200_SAO_EO_90 ? 0 : numVerVb, nH = typeIdx == B200_SAO_EO_0 ? 0 : numHorVb;
  for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) {
    const int x0 = x - dx[typeIdx], y0 = y - dy[typeIdx], x1 = x + dx[typeIdx], y1 = y + dy[typeIdx];
    const unsigned b0 = region_bit(x0, y0, width, height), b1 = region_bit(x1, y1, width, height);
    if ((b0 && !(avail & b0)) || (b1 && !(avail & b1))) continue;
    if (vb_disabled(x, y, nV, verVb, nH, horVb)) continue;
    const int c = src[y * srcStride + x];
    const int e = sgn(c - src[y0 * srcStride + x0]) + sgn(c - src[y1 * srcStride + x1]);
    dst[y * dstStride + x] = (int16_t)clip3(0, pmax, c + offset[2 + e]);
  }
}

void orc_sao_picture(const b200_geom* g, const int16_t* const src[3], int16_t* const dst[3], const b200_sao_ctu* ctus, const b200_vb* vb)
{
  const int ctu = g->ctuSize, ctusW = (g->width + ctu - 1) / ctu, ctusH = (g->height + ctu - 1) / ctu;
  const int nComp = g->chromaFormat ? 3 : 1;
  for (int c = 0; c < nComp; c++) {
    const int ph = c ? g->height >> 1 : g->height;
    memcpy(dst[c], src[c], (size_t)g->stride[c] * ph * sizeof(int16_t));     /* res == rec already holds the deblocked samples */
  }
  for (int cy = 0; cy < ctusH; cy++) for (int cx = 0; cx < ctusW; cx++) {
    const b200_sao_ctu* p = &ctus[cy * ctusW + cx];
    for (int c = 0; c < nComp; c++) {
      if (p->type[c] == B200_SAO_OFF) continue;
      const int sh = c ? 1 : 0;
      const int x0 = (cx * ctu) >> sh, y0 = (cy * ctu) >> sh;
      int w = ctu >> sh, h = ctu >> sh;
      const int pw = g->width >> sh, phh = g->height >> sh;
      if (x0 + w > pw) w = pw - x0;
      if (y0 + h > phh) h = phh - y0;
      int offs[32]; memset(offs, 0, sizeof(offs));
      if (p->type[c] == B200_SAO_BO) for (int i = 0; i < 4; i++) offs[(p->band[c] + i) & 31] = p->offset[c][i];
      else for (int i = 0; i < 5; i++) offs[i] = p->offset[c][i];
      int vv[3], hh[3], nV = 0, nH = 0;
      if (vb) {   /* isCrossedByVirtualBoundaries (UnitTools.cpp:3795): boundaries inside or ON the luma CTU area, made relative (:709-716) */
        const int lw = (cx * ctu + ctu > g->width ? g->width - cx * ctu : ctu), lh = (cy * ctu + ctu > g->height ? g->height - cy * ctu : ctu);
        for (int i = 0; i < vb->numVer; i++) if (vb->posX[i] >= cx * ctu && vb->posX[i] <= cx * ctu + lw) vv[nV++] = (vb->posX[i] >> sh) - x0;
        for (int i = 0; i < vb->numHor; i++) if (vb->posY[i] >= cy * ctu && vb->posY[i] <= cy * ctu + lh) hh[nH++] = (vb->posY[i] >> sh) - y0;
      }
      orc_sao_offset_block(g->bitDepth, p->type[c], offs, src[c] + (size_t)y0 * g->stride[c] + x0, dst[c] + (size_t)y0 * g->stride[c] + x0,
                           g->stride[c], g->stride[c], w, h, p->avail, nV, vv, nH, hh);
    }
  }
}
