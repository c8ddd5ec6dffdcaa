/* oracle/k1_residual.c — CPU restatement of K1 (dequant, LFNST, inverse transforms, TS/BDPCM/JCCR).
 * TEST INFRASTRUCTURE ONLY — see vvc_oracle.h. Pinned against oracle/_ref (tests/test_oracle_vs_ref.py). */
#include "vvc_oracle.h"
#include "../vvdec_b200/csrc/vvc_tables.h"
#include <string.h>
#include <stdlib.h>

static inline int32_t clip3(int32_t lo, int32_t hi, int32_t v) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- dequant: Quant.cpp:122-179 ------------------------------------------------------------ */
#define DEQUANT_BODY(QT)                                                                             \
  const int inputMinimum = -(inputMaximum + 1);                                                      \
  const int32_t transformMinimum = -(transformMaximum + 1);                                          \
  for (int y = 0; y <= maxY; y++) {                                                                  \
    for (int x = 0; x <= maxX; x++) {                                                                \
      const int n = y * width + x;                                                                   \
      const int32_t level = q[x + y * qStride];                                                      \
      if (!level) continue;                                                                          \
      const int sc = sl ? sl[n] * scale : scale;                                                     \
      const int32_t c = clip3(inputMinimum, inputMaximum, level);                                    \
      int32_t v;                                                                                     \
      if (rightShift > 0) v = (int32_t)((uint32_t)c * (uint32_t)sc + (1u << (rightShift - 1))) >> rightShift; \
      else                v = (int32_t)(((uint32_t)c * (uint32_t)sc) << (-rightShift));             \
      coef[n] = clip3(transformMinimum, transformMaximum, v);                                        \
    }                                                                                                \
  }

void orc_dequant(int width, int maxX, int maxY, int scale, const int32_t* sl, const int16_t* q, size_t qStride,
                 int32_t* coef, int rightShift, int inputMaximum, int32_t transformMaximum)
{ DEQUANT_BODY(int16_t) }

void orc_dequant32(int width, int maxX, int maxY, int scale, const int32_t* sl, const int32_t* q, size_t qStride,
                   int32_t* coef, int rightShift, int inputMaximum, int32_t transformMaximum)
{ DEQUANT_BODY(int32_t) }

/* ---- LFNST: TrQuant.cpp:79-106 ------------------------------------------------------------- */
void orc_inv_lfnst(const int32_t* src, int32_t* dst, unsigned set, unsigned index, unsigned size, int zeroOutSize)
{
  const int8_t* m = size > 4 ? &kLfnst8x8[(set * 2 + index) * 48 * 16] : &kLfnst4x4[(set * 2 + index) * 16 * 16];
  const int nOut = size > 4 ? 48 : 16;
  for (int j = 0; j < nOut; j++, m += 16) {
    int32_t acc = 0;
    for (int i = 0; i < zeroOutSize; i++) acc += src[i] * m[i];
    dst[j] = clip3(-32768, 32767, (acc + 64) >> 7);
  }
}

/* ---- 1-D inverse transform: TrQuant_EMT.cpp:103-121 (+ B2/B4 butterflies :126-223, which are the
 *      same integer sums reassociated; int32 wrap-around is modular so the results are identical) -- */
static const int16_t* tr_matrix(int trType, int n)
{
  if (trType == B200_TR_DCT2) {
    switch (n) { case 2: return kTrDCT2_2; case 4: return kTrDCT2_4; case 8: return kTrDCT2_8;
                 case 16: return kTrDCT2_16; case 32: return kTrDCT2_32; case 64: return kTrDCT2_64; }
  } else if (trType == B200_TR_DCT8) {
    switch (n) { case 4: return kTrDCT8_4; case 8: return kTrDCT8_8; case 16: return kTrDCT8_16; case 32: return kTrDCT8_32; }
  } else {
    switch (n) { case 4: return kTrDST7_4; case 8: return kTrDST7_8; case 16: return kTrDST7_16; case 32: return kTrDST7_32; }
  }
  abort();
}

void orc_inv_1d(int trType, int n, const int32_t* src, int32_t* dst, int shift, int line, int skipLine, int skipLine2,
                int clip, int32_t outMin, int32_t outMax)
{
  const int16_t* it = tr_matrix(trType, n);
  const int reducedLine = line - skipLine, cutoff = n - skipLine2;
  const int32_t rnd = 1 << (shift - 1);
  memset(dst, 0, sizeof(int32_t) * (size_t)line * n);
  for (int i = 0; i < reducedLine; i++)
    for (int j = 0; j < n; j++) {
      uint32_t acc = 0;
      for (int k = 0; k < cutoff; k++) acc += (uint32_t)src[k * line + i] * (uint32_t)(int32_t)it[k * n + j];
      int32_t v = (int32_t)acc;
      if (clip) v = clip3(outMin, outMax, (int32_t)((uint32_t)v + (uint32_t)rnd) >> shift);
      dst[i * n + j] = v;
    }
}

void orc_cpy_resi_clip(const int32_t* src, int16_t* dst, ptrdiff_t stride, unsigned w, unsigned h,
                       int32_t outMin, int32_t outMax, int32_t round, int32_t shift)
{
  for (unsigned y = 0; y < h; y++)
    for (unsigned x = 0; x < w; x++)
      dst[y * stride + x] = (int16_t)clip3(outMin, outMax, (src[y * w + x] + round) >> shift);
}

/* ---- one TU component: TrQuant.cpp:290 invTransformNxN ------------------------------------- */
void orc_tu_residual(const b200_tu* tu, int bitDepth, const int16_t* coefs, const int32_t* scaling,
                     int16_t* resi, ptrdiff_t stride)
{
  const int w = 1 << tu->log2w, h = 1 << tu->log2h;
  int maxX = tu->maxX, maxY = tu->maxY;
  const int32_t trMax = 32767, trMin = -32768;            /* maxLog2TrDynamicRange = 15 */
  const int inputMaximum = (1 << (tu->inBits - 1)) - 1;
  const int16_t* q = coefs + tu->coefOff;
  const size_t qStride = (size_t)maxX + 1;
  const int32_t* sl = (tu->flags & B200_TU_SCALING) ? scaling + tu->slOff : NULL;
  int32_t* dq  = (int32_t*)calloc((size_t)w * h, sizeof(int32_t));
  int32_t* tmp = (int32_t*)calloc((size_t)w * h, sizeof(int32_t));
  int32_t* blk = (int32_t*)calloc((size_t)w * h, sizeof(int32_t));

  if (tu->flags & (B200_TU_BDPCM_H | B200_TU_BDPCM_V)) {
    /* Quant.cpp:239 invResDPCM then DeQuantPCM over the whole block (:351-355) */
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        int32_t v = q[y * qStride + x];
        if ((tu->flags & B200_TU_BDPCM_H) && x > 0) v = clip3(trMin, trMax, dq[y * w + x - 1] + v);
        if ((tu->flags & B200_TU_BDPCM_V) && y > 0) v = clip3(trMin, trMax, dq[(y - 1) * w + x] + v);
        dq[y * w + x] = v;
      }
    orc_dequant32(w, w - 1, h - 1, tu->scale, sl, dq, (size_t)w, dq, tu->rightShift, inputMaximum, trMax);
  } else {
    orc_dequant(w, maxX, maxY, tu->scale, sl, q, qStride, dq, tu->rightShift, inputMaximum, trMax);
  }

  if (tu->lfnst && !(tu->flags & B200_TU_TS)) {
    /* TrQuant.cpp:201-288 xInvLfnst */
    const int idx = (tu->lfnst & 3) - 1, set = (tu->lfnst >> 2) & 3, transpose = (tu->lfnst >> 4) & 1;
    const int whge3 = w >= 8 && h >= 8;
    const int sb = whge3 ? 8 : 4;
    static const uint8_t sx[16] = {0,0,1,0,1,2,0,1,2,3,1,2,3,2,3,3}, sy[16] = {0,1,0,2,1,0,3,2,1,0,3,2,1,3,2,3};
    int32_t in[16], out[48];
    for (int i = 0; i < 16; i++) in[i] = dq[sy[i] * w + sx[i]];   /* g_coefTopLeftDiagScan8x8 / SCAN_GROUPED_4x4 first 16 */
    orc_inv_lfnst(in, out, set, idx, sb, ((w == 4 && h == 4) || (w == 8 && h == 8)) ? 8 : 16);
    const int32_t* o = out;
    if (transpose) {
      if (sb == 4) { for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) dq[y * w + x] = o[x * 4 + y]; }
      else {
        for (int y = 0; y < 8; y++) {
          for (int x = 0; x < 4; x++) dq[y * w + x] = o[x * 8 + y];
          if (y < 4) for (int x = 4; x < 8; x++) dq[y * w + x] = o[32 + (x - 4) * 4 + y];
        }
      }
    } else {
      for (int y = 0; y < sb; y++) { const int n = y < 4 ? sb : 4; for (int x = 0; x < n; x++) dq[y * w + x] = *o++; }
    }
    if (maxX < (w - 1 < 7 ? w - 1 : 7)) maxX = (w - 1 < 7 ? w - 1 : 7);
    if (maxY < (h - 1 < 7 ? h - 1 : 7)) maxY = (h - 1 < 7 ? h - 1 : 7);
  }

  if (tu->flags & B200_TU_TS) {
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) resi[y * stride + x] = (int16_t)dq[y * w + x];   /* TrQuant.cpp:489 */
  } else {
    /* TrQuant.cpp:410-485 xIT */
    const int trH = tu->trType & 3, trV = (tu->trType >> 2) & 3;
    const int shift1 = 7, shift2 = 20 - bitDepth;
    if (maxX == 0 && maxY == 0 && trH == B200_TR_DCT2 && trV == B200_TR_DCT2) {
      int32_t dc;
      if (w > 1 && h > 1) { dc = (dq[0] * 64 + (1 << (shift1 - 1))) >> shift1; dc = (dc * 64 + (1 << (shift2 - 1))) >> shift2; }
      else                { const int s = 21 - bitDepth; dc = (dq[0] * 64 + (1 << (s - 1))) >> s; }
      for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) resi[y * stride + x] = (int16_t)dc;
    } else {
      int skipW = (trH != B200_TR_DCT2 && w == 32) ? 16 : (w > 32 ? w - 32 : 0);
      int skipH = (trV != B200_TR_DCT2 && h == 32) ? 16 : (h > 32 ? h - 32 : 0);
      if (skipW < w - maxX - 1) skipW = w - maxX - 1;
      if (skipH < h - maxY - 1) skipH = h - maxY - 1;
      int shiftLast;
      if (w > 1 && h > 1) {
        orc_inv_1d(trV, h, dq,  tmp, shift1, w, skipW, skipH, 1, trMin, trMax);
        orc_inv_1d(trH, w, tmp, blk, shift2, h, 0,     skipW, 0, trMin, trMax);
        shiftLast = shift2;
      } else if (w == 1) { orc_inv_1d(trV, h, dq, blk, shift2 + 1, 1, 0, skipH, 0, trMin, trMax); shiftLast = shift2 + 1; }
      else               { orc_inv_1d(trH, w, dq, blk, shift2 + 1, 1, 0, skipW, 0, trMin, trMax); shiftLast = shift2 + 1; }
      orc_cpy_resi_clip(blk, resi, stride, w, h, trMin, trMax, 1 << (shiftLast - 1), shiftLast);
    }
  }
  free(dq); free(tmp); free(blk);
}

/* ---- list level: DecCu.cpp:536 reconstructResi (+ TrQuant.cpp:108 invTransformCbCr) -------- */
void orc_k1_residual(const b200_geom* g, int16_t* const planes[3], const b200_tu* tus, size_t numTus,
                     const int16_t* coefs, const int32_t* scaling, int mode)
{
  const int pmax = (1 << g->bitDepth) - 1;
  int16_t r0[64 * 64], r1[64 * 64];
  for (size_t t = 0; t < numTus; t++) {
    const b200_tu* tu = &tus[t];
    const int w = 1 << tu->log2w, h = 1 << tu->log2h;
    orc_tu_residual(tu, g->bitDepth, coefs, scaling, r0, w);
    int nOut = 1, comp1 = 0;
    if (tu->ict) {
      const int m = tu->ict;
      comp1 = tu->comp == 1 ? 2 : 1; nOut = 2;
      for (int i = 0; i < w * h; i++) {
        const int c = r0[i];
        r1[i] = (int16_t)((m == 2) ? c : (m == -2) ? -c : (m == 1 || m == 3) ? (c >> 1) : ((-c) >> 1));
      }
    }
    for (int o = 0; o < nOut; o++) {
      const int comp = o ? comp1 : tu->comp;
      const int16_t* r = o ? r1 : r0;
      int16_t* p = planes[comp] + (size_t)tu->y * g->stride[comp] + tu->x;
      for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        int16_t* d = &p[(size_t)y * g->stride[comp] + x];
        if (mode == 0) { int v = *d + r[y * w + x]; *d = (int16_t)(v < 0 ? 0 : v > pmax ? pmax : v); }   /* Buffer.cpp:83 recoCore */
        else *d = r[y * w + x];
      }
    }
  }
}
