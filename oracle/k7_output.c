/* k7_output.c — CPU restatement of the application layer's output formats.  TEST INFRASTRUCTURE ONLY (see vvc_oracle.h).
 * Follows /root/reference/source/App/vvdecapp/vvdecHelper.h:63-150 (_writeComponentToFile).  Pinned by tests/test_output_oracle_vs_ref.py. */
#include "vvc_oracle.h"
#include <stdlib.h>
#include <string.h>

void orc_pack_pyuv(const int16_t* src, ptrdiff_t stride, int w, int h, uint8_t* dst)
{
  for (int y = 0; y < h; y++, src += stride)
    for (int x = 0; x < w; x += 4) {
      const uint16_t* p = (const uint16_t*)src + x;
      const long long t = ((long long)p[0] << 0) + ((long long)p[1] << 10) + ((long long)p[2] << 20) + ((long long)p[3] << 30);
      for (int k = 0; k < 5; k++) *dst++ = (uint8_t)((t >> (8 * k)) & 0xff);
    }
}

void orc_narrow8(const int16_t* src, ptrdiff_t stride, int w, int h, int bitDepth, uint8_t* dst)
{
  for (int y = 0; y < h; y++, src += stride)
    for (int x = 0; x < w; x++) *dst++ = (uint8_t)(((const uint16_t*)src)[x] >> (bitDepth - 8));
}

/* Decoded-picture hash of one plane.  Follows /root/reference/source/Lib/CommonLib/PicYuvMD5.cpp: compCRC :100-136 (method 1: 16-bit
 * register 0xffff, polynomial 0x1021, message = per sample the low byte then — above 8 bit — the high byte, MSB first, then 16 zero
 * bits; digest = register hi, lo) and compChecksum :152-177 (method 2: sum of (byte ^ mask(x, y)) over the same bytes; digest big
 * endian).  Returns the digest length. */
int orc_plane_hash(int method, int bitDepth, const int16_t* src, ptrdiff_t stride, int w, int h, uint8_t* digest)
{
  if (method == 1) {
    uint32_t crc = 0xffff;
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        const int v = src[y * stride + x];
        for (int byte = 0; byte < (bitDepth > 8 ? 2 : 1); byte++)
          for (int bit = 7; bit >= 0; bit--) {
            const uint32_t msb = (crc >> 15) & 1, in = (uint32_t)(v >> (8 * byte + bit)) & 1;
            crc = (((crc << 1) + in) & 0xffff) ^ (msb * 0x1021);
          }
      }
    for (int bit = 0; bit < 16; bit++) { const uint32_t msb = (crc >> 15) & 1; crc = ((crc << 1) & 0xffff) ^ (msb * 0x1021); }
    digest[0] = (uint8_t)(crc >> 8); digest[1] = (uint8_t)crc;
    return 2;
  }
  uint32_t sum = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const uint8_t mask = (uint8_t)((x & 0xff) ^ (y & 0xff) ^ (x >> 8) ^ (y >> 8));
      const int v = src[y * stride + x];
      sum += (uint32_t)((v & 0xff) ^ mask);
      if (bitDepth > 8) sum += (uint32_t)((v >> 8) ^ mask);
    }
  digest[0] = (uint8_t)(sum >> 24); digest[1] = (uint8_t)(sum >> 16); digest[2] = (uint8_t)(sum >> 8); digest[3] = (uint8_t)sum;
  return 4;
}

/* ------------------------------------------------------------------------------------------------------------------------------
 * Film grain synthesis, the per-sample ("hardware") part.  Follows /root/reference/source/Lib/FilmGrain/FilmGrainImpl.cpp:
 * get_offset_y/u/v :85-127, add_grain_block :129-196 (overlap weights, offsets), make_grain_pattern :198-245 (pattern lookup by
 * intensity, vertical overlap with the block row above), scale_and_output :247-316 (horizontal smoothing across 16-sample block
 * borders, scaling, clipping to [0, 255 << bs]) and the block-seed walk of FilmGrain::add_grain_line (FilmGrain.cpp:836-867:
 * one prng step per 16-sample block, line seeds from prepareBlockSeeds :794).  The reference streams a line through a two-block
 * pipeline; restated here as a pure function of the sample position.  The pattern / LUT synthesis from the SEI (FilmGrain::init_sei,
 * the "firmware" part) is host code that runs once per SEI and is not part of this path: its tables are the input. */
static uint32_t fg_prng(uint32_t x) { const uint32_t s = ((x << 30) ^ (x << 2)) & 0x80000000u; return s | (x >> 1); }   /* FilmGrainImpl.h:70 */

static void fg_offsets(int c, uint32_t val, int csubx, int csuby, int* s, int* x, int* y)
{
  uint32_t bf;
  if (c == 0)      { *s = ((val >> 31) & 1) ? -1 : 1; bf = val & 0x3ff; *x = (int)((bf * 13) >> 10) * 4; bf = (val >> 14) & 0x3ff; *y = (int)((bf * 12) >> 10) * 4; }
  else if (c == 1) { *s = ((val >> 2) & 1) ? -1 : 1; bf = (val >> 10) & 0x3ff; *x = (int)((bf * 13) >> 10) * (4 / csubx);
                     bf = ((val >> 24) & 0xff) | ((val << 8) & 0x300); *y = (int)((bf * 12) >> 10) * (4 / csuby); }
  else             { *s = ((val >> 15) & 1) ? -1 : 1; bf = (val >> 20) & 0x3ff; *x = (int)((bf * 13) >> 10) * (4 / csubx);
                     bf = (val >> 4) & 0x3ff; *y = (int)((bf * 12) >> 10) * (4 / csuby); }
  *x &= 0xff; *y &= 0xff;
}

typedef struct { const int16_t* src; ptrdiff_t stride; int c, subx, suby, bs; const int8_t* pattern; const uint8_t *sLUT, *pLUT; const uint32_t* seeds; int nbx; } fg_ctx;

/* grain of component sample (xs, ys) before the horizontal smoothing: make_grain_pattern for the block that holds it */
static int fg_grain(const fg_ctx* f, int xs, int ys)
{
  const int y = ys * f->suby, bw = 16 / f->subx, bx = xs / bw, i = xs - bx * bw, by = y >> 4, j = y & 15;
  int oc1 = 0, oc2 = 0;
  if (y > 15 && j == 0) { oc1 = f->suby > 1 ? 20 : 12; oc2 = f->suby > 1 ? 20 : 24; }
  else if (y > 15 && j == 1) { oc1 = 24; oc2 = 12; }
  int s, ox, oy;
  fg_offsets(f->c, f->seeds[by * f->nbx + bx], f->subx, f->suby, &s, &ox, &oy);
  oy += j / f->suby;
  const int intensity = (((const uint16_t*)f->src)[ys * f->stride + xs] >> f->bs) & 0xff;
  const int pi = f->pLUT[f->c * 256 + intensity] >> 4;
  const int8_t* pat = f->pattern + ((f->c ? 1 : 0) * 8 + pi) * 4096;
  int P = pat[oy * 64 + ox + i] * s;
  if (oc1) {
    int sUp, oxUp, oyUp;
    fg_offsets(f->c, f->seeds[(by - 1) * f->nbx + bx], f->subx, f->suby, &sUp, &oxUp, &oyUp);
    oyUp += (16 + j) / f->suby;
    P = (P * oc1 + pat[oyUp * 64 + oxUp + i] * oc2 * sUp + 16) >> 5;
  }
  return P;
}

/* pattern [2][8][64][64], sLUT / pLUT [3][256], lineSeeds [(h+15)/16]; planes in place; 4:2:0 or 4:0:0 (planes[1] == NULL) */
void orc_film_grain(int16_t* const planes[3], const ptrdiff_t strides[3], int w, int h, int bitDepth, const int8_t* pattern, const uint8_t* sLUT,
                    const uint8_t* pLUT, const uint32_t* lineSeeds, int scaleShift, const uint8_t compPresent[3])
{
  const int nbx = (w + 15) / 16, nby = (h + 15) / 16, bs = bitDepth - 8;
  uint32_t* seeds = (uint32_t*)malloc(sizeof(uint32_t) * nbx * nby);
  for (int by = 0; by < nby; by++) { uint32_t r = lineSeeds[by]; for (int bx = 0; bx < nbx; bx++) { seeds[by * nbx + bx] = r; r = fg_prng(r); } }
  for (int c = 0; c < 3; c++) {
    if (!planes[c] || !compPresent[c]) continue;
    const int sub = c ? 2 : 1, cw = w / sub, ch = h / sub, bw = 16 / sub;
    fg_ctx f = {planes[c], strides[c], c, sub, sub, bs, pattern, sLUT, pLUT, seeds, nbx};
    int16_t* out = (int16_t*)malloc(sizeof(int16_t) * cw * ch);
    for (int y = 0; y < ch; y++)
      for (int x = 0; x < cw; x++) {
        const int bx = x / bw, i = x - bx * bw;
        int g = fg_grain(&f, x, y);
        /* smoothing across the border between two blocks of the line (scale_and_output :262-274): both samples next to it */
        if ((i == 0 && bx > 0) || (i == bw - 1 && bx + 1 < nbx)) g = (fg_grain(&f, x - 1, y) + 3 * g + fg_grain(&f, x + 1, y) + 2) >> 2;
        const int v = ((const uint16_t*)planes[c])[y * strides[c] + x], scale = sLUT[c * 256 + ((v >> bs) & 0xff)];
        const int add = (scale * (int16_t)g + (1 << (scaleShift - 1))) >> scaleShift;
        int o = v + add; if (o < 0) o = 0; if (o > (255 << bs)) o = 255 << bs;
        out[y * cw + x] = (int16_t)o;
      }
    for (int y = 0; y < ch; y++) memcpy(planes[c] + y * strides[c], out + y * cw, sizeof(int16_t) * cw);
    free(out);
  }
  free(seeds);
}
