// k4_sao.cu — K4: sample adaptive offset, src (deblocked) -> dst, one thread per 4 consecutive samples.
//
// Replaces (reference, source/Lib/CommonLib/SampleAdaptiveOffset.cpp): offsetBlock_core :64-349, SAOProcessCTU :522,
// offsetCTU :661, isProcessDisabled :817.  The CPU version walks lines with sign caches; per sample it is
//   EO: dst = clip(c + offset[2 + sgn(c-n0) + sgn(c-n1)]) if both neighbours along the class direction lie inside the
//       CTU or in an available neighbouring CTU and the sample is not adjacent to a virtual boundary; else dst = c
//   BO: dst = clip(c + offset[c >> (bd-5)])
// which is what each thread evaluates.  HBM traffic: S*2 B read (neighbour rows hit L1/L2) + S*2 B written + 24 B/CTU.
#include "common.cuh"

namespace b200 {

struct SaoParams {
  const int16_t* src[3]; int16_t* dst[3]; int stride[3];
  int W, H, bitDepth, ctuSize, ctusW, chroma;
  const b200_sao_ctu* ctus;
  b200_vb vb;
};

__device__ __forceinline__ int sgn(int v) { return (v > 0) - (v < 0); }

__global__ void __launch_bounds__(256) sao_kernel(const SaoParams P)
{
  const int c = blockIdx.z;
  const int sh = c ? 1 : 0;
  const int pw = P.W >> sh, ph = P.H >> sh;
  const int x = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= pw || y >= ph) return;
  const int stride = c == 0 ? P.stride[0] : c == 1 ? P.stride[1] : P.stride[2];
  const int16_t* s = c == 0 ? P.src[0] : c == 1 ? P.src[1] : P.src[2];
  const int cs = P.ctuSize >> sh;                       // CTU size in this plane
  const int cxi = x / cs, cyi = y / cs;
  // the 24-byte CTU record as six 32-bit words (one L1 line for the whole warp)
  const uint32_t* cpw = reinterpret_cast<const uint32_t*>(P.ctus + cyi * P.ctusW + cxi);
  const uint32_t w0 = __ldg(cpw), w1 = __ldg(cpw + 1);
  const int type = c == 0 ? (w0 & 0xff) : c == 1 ? ((w0 >> 8) & 0xff) : ((w0 >> 16) & 0xff);
  uint2* dptr = reinterpret_cast<uint2*>((c == 0 ? P.dst[0] : c == 1 ? P.dst[1] : P.dst[2]) + (size_t)y * stride + x);
  const uint2 ctr = *reinterpret_cast<const uint2*>(s + (size_t)y * stride + x);
  if (type == B200_SAO_OFF) { *dptr = ctr; return; }
  // offsets: bytes 6..20 of the record = offset[3][5]
  const uint32_t w2 = __ldg(cpw + 2), w3 = __ldg(cpw + 3), w4 = __ldg(cpw + 4), w5 = __ldg(cpw + 5);
  auto rec_byte = [&](int b) -> int { const uint32_t w = b < 8 ? w1 : b < 12 ? w2 : b < 16 ? w3 : b < 20 ? w4 : w5; return (int)(int8_t)((w >> ((b & 3) * 8)) & 0xff); };
  int off[5];
#pragma unroll
  for (int k = 0; k < 5; k++) off[k] = rec_byte(6 + c * 5 + k);
  const int band = c == 0 ? (w0 >> 24) : c == 1 ? (w1 & 0xff) : ((w1 >> 8) & 0xff);
  const unsigned avail = (w5 >> 8) & 0xff;                  // byte 21
  int v[4] = { (int)(int16_t)(ctr.x & 0xffff), (int)(int16_t)(ctr.x >> 16), (int)(int16_t)(ctr.y & 0xffff), (int)(int16_t)(ctr.y >> 16) };
  int r[4] = { v[0], v[1], v[2], v[3] };
  const int pmax = (1 << P.bitDepth) - 1;
  if (type == B200_SAO_BO) {
    const int shiftBits = P.bitDepth - 5;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int k = ((v[i] >> shiftBits) - band) & 31;
      if (k < 4) r[i] = clip3(0, pmax, v[i] + (k == 0 ? off[0] : k == 1 ? off[1] : k == 2 ? off[2] : off[3]));
    }
  } else {
    // neighbour offsets of the class: EO_0 (-1,0)/(+1,0); EO_90 (0,-1)/(0,+1); EO_135 (-1,-1)/(+1,+1); EO_45 (+1,-1)/(-1,+1)
    const int dx = type == B200_SAO_EO_90 ? 0 : (type == B200_SAO_EO_45 ? -1 : 1);
    const int dy = type == B200_SAO_EO_0 ? 0 : 1;
    const int x0c = cxi * cs, y0c = cyi * cs;
    const int w = min(cs, pw - x0c), h = min(cs, ph - y0c);
    const int nV = type == B200_SAO_EO_90 ? 0 : P.vb.numVer, nH = type == B200_SAO_EO_0 ? 0 : P.vb.numHor;
    const int ly = y - y0c;
    // rows of the two neighbours; interior threads (not on the CTU border) skip the availability logic
    const bool interior = ly > 0 && ly < h - 1 && (x - x0c) > 0 && (x - x0c) + 4 < w && nV == 0 && nH == 0;
    // neighbour samples of the 4 outputs, fetched as two 8-byte vectors + at most two scalars (rows / columns clamped: clamped
    // values are only ever used by samples that the availability test rejects)
    int na[4], nb[4];
    {
      const int16_t* rowA = s + (size_t)max(y - dy, 0) * stride;
      const int16_t* rowB = s + (size_t)min(y + dy, ph - 1) * stride;
      const int xl = max(x - 1, 0), xr = min(x + 4, pw - 1);
      if (dy == 0) {
        const int l = rowA[xl], r4 = rowA[xr];
        na[0] = l; na[1] = v[0]; na[2] = v[1]; na[3] = v[2]; nb[0] = v[1]; nb[1] = v[2]; nb[2] = v[3]; nb[3] = r4;
      } else {
        const uint2 ua = *reinterpret_cast<const uint2*>(rowA + x), ub = *reinterpret_cast<const uint2*>(rowB + x);
        const int A[4] = { (int)(int16_t)(ua.x & 0xffff), (int)(int16_t)(ua.x >> 16), (int)(int16_t)(ua.y & 0xffff), (int)(int16_t)(ua.y >> 16) };
        const int B[4] = { (int)(int16_t)(ub.x & 0xffff), (int)(int16_t)(ub.x >> 16), (int)(int16_t)(ub.y & 0xffff), (int)(int16_t)(ub.y >> 16) };
        if (dx == 0)      { na[0] = A[0]; na[1] = A[1]; na[2] = A[2]; na[3] = A[3]; nb[0] = B[0]; nb[1] = B[1]; nb[2] = B[2]; nb[3] = B[3]; }
        else if (dx == 1) { na[0] = rowA[xl]; na[1] = A[0]; na[2] = A[1]; na[3] = A[2]; nb[0] = B[1]; nb[1] = B[2]; nb[2] = B[3]; nb[3] = rowB[xr]; }
        else              { na[0] = A[1]; na[1] = A[2]; na[2] = A[3]; na[3] = rowA[xr]; nb[0] = rowB[xl]; nb[1] = B[0]; nb[2] = B[1]; nb[3] = B[2]; }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int gx = x + i;
      bool ok = true;
      if (!interior) {
        const int lx = gx - x0c;
#pragma unroll
        for (int n = 0; n < 2; n++) {
          const int nx = lx + (n ? dx : -dx), ny = ly + (n ? dy : -dy);
          const bool l = nx < 0, rr = nx >= w, a = ny < 0, b = ny >= h;
          unsigned bit = 0;
          if (a)      bit = l ? B200_AVAIL_AL : rr ? B200_AVAIL_AR : B200_AVAIL_A;
          else if (b) bit = l ? B200_AVAIL_BL : rr ? B200_AVAIL_BR : B200_AVAIL_B;
          else        bit = l ? B200_AVAIL_L : rr ? B200_AVAIL_R : 0;
          if (bit && !(avail & bit)) ok = false;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) { const int p = P.vb.posX[k] >> sh; if (k < nV && (gx == p || gx == p - 1)) ok = false; }
#pragma unroll
        for (int k = 0; k < 3; k++) { const int p = P.vb.posY[k] >> sh; if (k < nH && (y == p || y == p - 1)) ok = false; }
      }
      if (!ok) continue;
      const int e = sgn(v[i] - na[i]) + sgn(v[i] - nb[i]);
      r[i] = clip3(0, pmax, v[i] + (e == -2 ? off[0] : e == -1 ? off[1] : e == 0 ? off[2] : e == 1 ? off[3] : off[4]));
    }
  }
  uint2 o;
  o.x = (unsigned)(r[0] & 0xffff) | ((unsigned)r[1] << 16);
  o.y = (unsigned)(r[2] & 0xffff) | ((unsigned)r[3] << 16);
  *dptr = o;
}

int launch_sao(const SaoLaunch& L, cudaStream_t s, KProf* prof)
{
  SaoParams P;
  for (int c = 0; c < 3; c++) { P.src[c] = L.src.p[c]; P.dst[c] = L.dst.p[c]; P.stride[c] = L.src.stride[c]; }
  P.W = L.geom.width; P.H = L.geom.height; P.bitDepth = L.geom.bitDepth; P.ctuSize = L.geom.ctuSize;
  P.ctusW = (P.W + P.ctuSize - 1) / P.ctuSize; P.chroma = L.geom.chromaFormat == 1;
  P.ctus = L.ctus; P.vb = L.vb;
  dim3 blk(32, 8), grd((P.W / 4 + 31) / 32, (P.H + 7) / 8, P.chroma ? 3 : 1);
  if (prof) prof->begin(B200_KF_SAO, s);
  sao_kernel<<<grd, blk, 0, s>>>(P);
  B200_CUDA(cudaGetLastError());
  if (prof) prof->end(B200_KF_SAO, s);
  return 0;
}

}  // namespace b200
