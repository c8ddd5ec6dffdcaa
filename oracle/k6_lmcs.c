/* k6_lmcs.c — CPU restatement of the LMCS steps on the inter reconstruction path.  TEST INFRASTRUCTURE ONLY (see vvc_oracle.h).
 * Follows /root/reference/source/Lib: CommonLib/Reshape.cpp (calculateChromaAdjVpduNei :192-277, getPWLIdxInv :283-291, rspCtuBcw :377-408,
 * rspBufFwd :410-413), CommonLib/Buffer.cpp (applyLutCore :200-215, rspFwdCore :321-339, scaleSignal :412-439),
 * DecoderLib/DecCu.cpp (predAndReco :458-476, finishLMCSAndReco :483-533).  Pinned against the compiled reference by
 * tests/test_lmcs_oracle_vs_ref.py. */
#include "vvc_oracle.h"
#include <stdlib.h>
#include <string.h>

static int ilog2(int v) { int r = 0; while ((1 << (r + 1)) <= v) r++; return r; }
static int clip_bd(int v, int bd) { const int m = (1 << bd) - 1; return v < 0 ? 0 : v > m ? m : v; }

void orc_lmcs_fwd_block(int16_t* ptr, ptrdiff_t stride, int w, int h, int bitDepth, const b200_lmcs* L)
{
  const int shift = ilog2(L->orgCW);
  for (int y = 0; y < h; y++, ptr += stride)
    for (int x = 0; x < w; x++) {
      const int v = ptr[x], idxY = v >> shift;
      ptr[x] = (int16_t)clip_bd(L->reshapePivot[idxY] + ((L->fwdScaleCoef[idxY] * (v - L->inputPivot[idxY]) + (1 << 10)) >> 11), bitDepth);
    }
}

void orc_lmcs_fwd_pus(const b200_geom* g, int16_t* luma, const b200_pu* pus, size_t numPus, const b200_lmcs* L)
{
  for (size_t i = 0; i < numPus; i++)
    orc_lmcs_fwd_block(luma + (size_t)pus[i].y * g->stride[0] + pus[i].x, g->stride[0], pus[i].w, pus[i].h, g->bitDepth, L);
}

static int pwl_idx_inv(const b200_lmcs* L, int lumaVal)   /* Reshape.cpp:283 */
{
  int idxS;
  for (idxS = L->minBinIdx; idxS <= L->maxBinIdx; idxS++)
    if (lumaVal < L->reshapePivot[idxS + 1]) break;
  return idxS < 15 ? idxS : 15;
}

int orc_lmcs_vpdu_scale(const b200_geom* g, const int16_t* luma, const b200_lmcs* L, const b200_lmcs_vpdu* v)
{
  const int numNeighbor = g->ctuSize < 64 ? g->ctuSize : 64, numNeighborLog = ilog2(numNeighbor);
  const int xPos = v->x, yPos = v->y, picW = g->width, picH = g->height;
  const int16_t* rec = luma + (size_t)yPos * g->stride[0] + xPos;
  int recLuma = 0, pelnum = 0, lumaValue;
  if (v->availLeft)
    for (int i = 0; i < numNeighbor; i++) { const int k = (yPos + i) >= picH ? (picH - yPos - 1) : i; recLuma += rec[-1 + (ptrdiff_t)k * g->stride[0]]; pelnum++; }
  if (v->availAbove)
    for (int i = 0; i < numNeighbor; i++) { const int k = (xPos + i) >= picW ? (picW - xPos - 1) : i; recLuma += rec[-(ptrdiff_t)g->stride[0] + k]; pelnum++; }
  if (pelnum == numNeighbor) lumaValue = (recLuma + (1 << (numNeighborLog - 1))) >> numNeighborLog;
  else if (pelnum == (numNeighbor << 1)) lumaValue = (recLuma + (1 << numNeighborLog)) >> (numNeighborLog + 1);
  else lumaValue = 1 << (g->bitDepth - 1);
  return L->chromaAdjHelpLUT[pwl_idx_inv(L, lumaValue)];
}

int orc_lmcs_scale_resi(int r, int scale, int bitDepth)
{
  const int maxAbs = (1 << bitDepth) - 1;
  r = r < -maxAbs - 1 ? -maxAbs - 1 : r > maxAbs ? maxAbs : r;
  const int sign = r >= 0 ? 1 : -1, absval = sign * r;
  int val = sign * ((absval * scale + (1 << 10)) >> 11);
  return val < -32768 ? -32768 : val > 32767 ? 32767 : val;
}

void orc_k1_residual_lmcs(const b200_geom* g, int16_t* const planes[3], const b200_tu* tus, size_t numTus,
                          const int16_t* coefs, const int32_t* scaling, const b200_lmcs* L)
{
  const int pmax = (1 << g->bitDepth) - 1;
  /* luma TUs first: their reconstruction (mapped domain) feeds the chroma scale of every VPDU */
  for (size_t t = 0; t < numTus; t++) if (tus[t].comp == 0) orc_k1_residual(g, planes, &tus[t], 1, coefs, scaling, 0);
  if (!L->chromaAdj) {
    for (size_t t = 0; t < numTus; t++) if (tus[t].comp != 0) orc_k1_residual(g, planes, &tus[t], 1, coefs, scaling, 0);
    return;
  }
  const int vs = g->ctuSize == 128 ? 64 : g->ctuSize, vl = ilog2(vs), vW = (g->width + vs - 1) / vs, vH = (g->height + vs - 1) / vs;
  int* scale = (int*)malloc(sizeof(int) * (size_t)vW * vH);
  for (int i = 0; i < vW * vH; i++) scale[i] = orc_lmcs_vpdu_scale(g, planes[0], L, &L->vpdus[i]);
  int16_t r0[64 * 64], r1[64 * 64];
  for (size_t t = 0; t < numTus; t++) {
    const b200_tu* tu = &tus[t];
    if (tu->comp == 0) continue;
    const int w = 1 << tu->log2w, h = 1 << tu->log2h;
    orc_tu_residual(tu, g->bitDepth, coefs, scaling, r0, w);
    int nOut = 1, comp1 = 0;
    if (tu->ict) {
      const int m = tu->ict;
      comp1 = tu->comp == 1 ? 2 : 1; nOut = 2;
      for (int i = 0; i < w * h; i++) { const int c = r0[i]; r1[i] = (int16_t)((m == 2) ? c : (m == -2) ? -c : (m == 1 || m == 3) ? (c >> 1) : ((-c) >> 1)); }
    }
    /* the TU's luma block starts at twice the chroma position (4:2:0); DecCu.cpp:504 passes that block, Reshape.cpp:200-210 masks it to the VPDU */
    const int sc = scale[((tu->y * 2) >> vl) * vW + ((tu->x * 2) >> vl)];
    const int doScale = w * h > 4;                         /* DecCu.cpp:506: blocks[compID].area() > 4 */
    for (int o = 0; o < nOut; o++) {
      const int comp = o ? comp1 : tu->comp;
      const int16_t* r = o ? r1 : r0;
      int16_t* p = planes[comp] + (size_t)tu->y * g->stride[comp] + tu->x;
      for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        int16_t* d = &p[(size_t)y * g->stride[comp] + x];
        const int rr = doScale ? orc_lmcs_scale_resi(r[y * w + x], sc, g->bitDepth) : r[y * w + x];
        const int v = *d + rr;
        *d = (int16_t)(v < 0 ? 0 : v > pmax ? pmax : v);
      }
    }
  }
  free(scale);
}

void orc_lmcs_inv_plane(const b200_geom* g, int16_t* luma, const b200_lmcs* L)
{
  for (int y = 0; y < g->height; y++) for (int x = 0; x < g->width; x++) { int16_t* p = &luma[(size_t)y * g->stride[0] + x]; *p = L->invLUT[*p]; }
}

/* ---- LMCS together with intra / CIIP CUs (DecCu.cpp:316-403 intra branch with doChrScale :619-633, :483 finishLMCSAndReco; the chroma scale of a VPDU
 * comes from its reconstructed luma neighbourhood, intra blocks included, so the chain is: luma TUs -> luma intra blocks -> scales -> chroma TUs ->
 * chroma intra blocks).  orc_lmcs_vpdu_scales: the scale of every VPDU from the luma plane as it stands. */
void orc_lmcs_vpdu_scales(const b200_geom* g, const int16_t* luma, const b200_lmcs* L, int32_t* scale)
{
  const int vs = g->ctuSize == 128 ? 64 : g->ctuSize, vW = (g->width + vs - 1) / vs, vH = (g->height + vs - 1) / vs;
  for (int i = 0; i < vW * vH; i++) scale[i] = orc_lmcs_vpdu_scale(g, luma, L, &L->vpdus[i]);
}

/* K1 over the TUs of one channel (compSel 1 luma, 2 chroma, 0 all): TUs flagged B200_TU_RESI leave their residual in the resi planes (K6 adds it),
 * the others are reconstructed into the planes; chroma residuals of blocks with more than 4 samples are scaled by their VPDU's scale (scale != NULL). */
void orc_k1_residual_sel(const b200_geom* g, int16_t* const planes[3], int16_t* const resi[3], const b200_tu* tus, size_t numTus,
                         const int16_t* coefs, const int32_t* scaling, int compSel, const int32_t* scale)
{
  const int pmax = (1 << g->bitDepth) - 1;
  const int vs = g->ctuSize == 128 ? 64 : g->ctuSize, vl = ilog2(vs), vW = (g->width + vs - 1) / vs;
  int16_t r0[64 * 64], r1[64 * 64];
  for (size_t t = 0; t < numTus; t++) {
    const b200_tu* tu = &tus[t];
    if ((compSel == 1 && tu->comp != 0) || (compSel == 2 && tu->comp == 0)) continue;
    const int w = 1 << tu->log2w, h = 1 << tu->log2h;
    orc_tu_residual(tu, g->bitDepth, coefs, scaling, r0, w);
    int nOut = 1, comp1 = 0;
    if (tu->ict) {
      const int m = tu->ict;
      comp1 = tu->comp == 1 ? 2 : 1; nOut = 2;
      for (int i = 0; i < w * h; i++) { const int c = r0[i]; r1[i] = (int16_t)((m == 2) ? c : (m == -2) ? -c : (m == 1 || m == 3) ? (c >> 1) : ((-c) >> 1)); }
    }
    const int doScale = scale && tu->comp != 0 && w * h > 4;
    const int sc = doScale ? scale[((tu->y * 2) >> vl) * vW + ((tu->x * 2) >> vl)] : 0;
    const int toResi = (tu->flags & B200_TU_RESI) && resi;
    for (int o = 0; o < nOut; o++) {
      const int comp = o ? comp1 : tu->comp;
      const int16_t* r = o ? r1 : r0;
      int16_t* p = (toResi ? resi[comp] : planes[comp]) + (size_t)tu->y * g->stride[comp] + tu->x;
      for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        int16_t* d = &p[(size_t)y * g->stride[comp] + x];
        const int rr = doScale ? orc_lmcs_scale_resi(r[y * w + x], sc, g->bitDepth) : r[y * w + x];
        if (toResi) *d = (int16_t)rr;
        else { const int v = *d + rr; *d = (int16_t)(v < 0 ? 0 : v > pmax ? pmax : v); }
      }
    }
  }
}
