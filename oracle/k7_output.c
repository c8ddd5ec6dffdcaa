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
