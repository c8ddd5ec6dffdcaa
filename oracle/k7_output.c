/* k7_output.c — CPU restatement of the application layer's output formats.  TEST INFRASTRUCTURE ONLY (see vvc_oracle.h).
 * Follows /root/reference/source/App/vvdecapp/vvdecHelper.h:63-150 (_writeComponentToFile).  Pinned by tests/test_output_oracle_vs_ref.py. */
#include "vvc_oracle.h"

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
