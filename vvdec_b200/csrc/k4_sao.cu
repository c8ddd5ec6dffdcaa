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
  const int stride = P.stride[c];
  const int16_t* s = P.src[c];
  const int cs = P.ctuSize >> sh;                       // CTU size in this plane
  const int cxi = x / cs, cyi = y / cs;
  const b200_sao_ctu& cp = P.ctus[cyi * P.ctusW + cxi];
  const int type = cp.type[c];
  const uint2 ctr = *reinterpret_cast<const uint2*>(s + (size_t)y * stride + x);
  int v[4] = { (int)(int16_t)(ctr.x & 0xffff), (int)(int16_t)(ctr.x >> 16), (int)(int16_t)(ctr.y & 0xffff), (int)(int16_t)(ctr.y >> 16) };
  int r[4] = { v[0], v[1], v[2], v[3] };
  const int pmax = (1 << P.bitDepth) - 1;
  if (type == B200_SAO_BO) {
    const int shiftBits = P.bitDepth - 5, band = cp.band[c];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int k = ((v[i] >> shiftBits) - band) & 31;
      if (k < 4) r[i] = clip3(0, pmax, v[i] + cp.offset[c][k]);
    }
  } else if (type != B200_SAO_OFF) {
    // neighbour offsets of the class: EO_0 (-1,0)/(+1,0); EO_90 (0,-1)/(0,+1); EO_135 (-1,-1)/(+1,+1); EO_45 (+1,-1)/(-1,+1)
    const int dx = type == B200_SAO_EO_90 ? 0 : (type == B200_SAO_EO_45 ? -1 : 1);
    const int dy = type == B200_SAO_EO_0 ? 0 : 1;
    const int x0c = cxi * cs, y0c = cyi * cs;
    const int w = min(cs, pw - x0c), h = min(cs, ph - y0c);
    const unsigned avail = cp.avail;
    const int nV = type == B200_SAO_EO_90 ? 0 : P.vb.numVer, nH = type == B200_SAO_EO_0 ? 0 : P.vb.numHor;
    const int ly = y - y0c;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int gx = x + i, lx = gx - x0c;
      // region of the two neighbours relative to the CTU -> availability bit (0: inside)
      bool ok = true;
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
      for (int k = 0; k < nV; k++) { const int p = P.vb.posX[k] >> sh; if (gx == p || gx == p - 1) ok = false; }
      for (int k = 0; k < nH; k++) { const int p = P.vb.posY[k] >> sh; if (y == p || y == p - 1) ok = false; }
      if (!ok) continue;
      const int n0 = s[(size_t)(y - dy) * stride + gx - dx], n1 = s[(size_t)(y + dy) * stride + gx + dx];
      const int e = sgn(v[i] - n0) + sgn(v[i] - n1);
      r[i] = clip3(0, pmax, v[i] + cp.offset[c][2 + e]);
    }
  }
  uint2 o;
  o.x = (unsigned)(r[0] & 0xffff) | ((unsigned)r[1] << 16);
  o.y = (unsigned)(r[2] & 0xffff) | ((unsigned)r[3] << 16);
  *reinterpret_cast<uint2*>(P.dst[c] + (size_t)y * stride + x) = o;
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
