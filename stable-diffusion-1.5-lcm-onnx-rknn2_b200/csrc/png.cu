// PNG files assembled ON THE DEVICE from the decoder's u8 images.
//
// The reference ends every job with `img.save(buf, format="PNG")` on the host (reference
// `backends/cuda_worker.py:234-239`): 70-100 ms of zlib per 512x512 image and core, i.e. ~350
// images/s on a 32-core box against > 1000 images/s of kernels on 8 GPUs.  The job contract is "PNG
// bytes, same seed => same bytes" (reference `tests/test_sdxl_worker.py:139-198`), not "PIL's bytes", so
// with B200_PNG=gpu the worker takes the finished file from here: 8-bit RGB, filter type 0 on every
// scanline, a zlib stream of STORED deflate blocks (RFC 1951 §3.2.4, no entropy coding), Adler-32 and
// CRC-32 computed by the kernels below.  Any PNG reader decodes it to exactly the decoder's pixels.
//
//   offset  0  signature (8)
//           8  IHDR chunk (25): 13 | "IHDR" | w | h | 8 | 2 | 0 | 0 | 0 | crc
//          33  IDAT chunk: Z | "IDAT" | 78 01 | nblk x { final? | LEN | ~LEN | <= 65535 raw bytes } | adler32 | crc
//              raw = h scanlines of (1 filter byte 0 + 3w pixel bytes)
//        45+Z  IEND chunk (12)
#include "common.cuh"
#include "dreamlab_b200.h"

namespace dl {

constexpr int PNG_BLOCK = 65535;          // stored-block payload limit
constexpr int PNG_SEG = 4096;             // CRC segment one thread walks
constexpr int PNG_IDAT = 33;              // file offset of the IDAT chunk

struct PngGeom {
  int h, w;
  long long raw;        // h * (1 + 3w)
  int nblk;             // stored blocks
  long long z;          // zlib stream bytes = 2 + raw + 5 nblk + 4
  long long total;      // file bytes
  uint32_t ihdr_crc;
  uint32_t op_seg[32];  // GF(2) operator: CRC register advanced over PNG_SEG zero bytes
  uint32_t op_tail[32]; // ... over the length of the last (shorter) segment
};

__constant__ uint32_t c_crc_table[256];

static uint32_t h_crc_table[256];
static bool h_crc_ready = false;

static void host_crc_init() {
  if (h_crc_ready) return;
  for (uint32_t n = 0; n < 256; ++n) {
    uint32_t c = n;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    h_crc_table[n] = c;
  }
  h_crc_ready = true;
}
static uint32_t host_crc(const uint8_t* p, size_t n) {
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; ++i) c = h_crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
// zlib's crc32_combine in operator form: crc(A || B) = M^{8 |B|} crc(A) xor crc(B), M = one zero bit
static uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
  uint32_t sum = 0;
  for (int i = 0; vec; vec >>= 1, ++i)
    if (vec & 1) sum ^= mat[i];
  return sum;
}
static void gf2_square(uint32_t* sq, const uint32_t* mat) {
  for (int n = 0; n < 32; ++n) sq[n] = gf2_times(mat, mat[n]);
}
static void crc_zero_operator(uint32_t* op, unsigned long long nbytes) {
  uint32_t even[32], odd[32];
  odd[0] = 0xEDB88320u;                       // one zero bit
  uint32_t row = 1;
  for (int n = 1; n < 32; ++n) { odd[n] = row; row <<= 1; }
  gf2_square(even, odd);                      // two bits
  gf2_square(odd, even);                      // four bits
  for (int n = 0; n < 32; ++n) op[n] = 1u << n;           // identity
  uint32_t tmp[32];
  // odd = 4 bits; each squaring doubles: first square -> 8 bits = one byte
  unsigned long long len = nbytes;
  uint32_t* cur = odd;
  uint32_t* nxt = even;
  while (len) {
    gf2_square(nxt, cur);                     // operator for the next power of two bytes
    uint32_t* t = cur; cur = nxt; nxt = t;
    if (len & 1) {
      for (int n = 0; n < 32; ++n) tmp[n] = gf2_times(cur, op[n]);
      for (int n = 0; n < 32; ++n) op[n] = tmp[n];
    }
    len >>= 1;
  }
}

static PngGeom png_geom(int h, int w) {
  PngGeom g;
  memset(&g, 0, sizeof(g));
  g.h = h; g.w = w;
  g.raw = (long long)h * (1 + 3LL * w);
  g.nblk = (int)((g.raw + PNG_BLOCK - 1) / PNG_BLOCK);
  g.z = 2 + g.raw + 5LL * g.nblk + 4;
  g.total = 8 + 25 + (12 + g.z) + 12;
  return g;
}

__device__ __forceinline__ uint32_t dev_gf2_times(const uint32_t* mat, uint32_t vec) {
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if ((vec >> i) & 1) sum ^= mat[i];
  return sum;
}

// file byte at offset `off` for everything except the Adler-32 and the IDAT CRC fields (written later)
__device__ __forceinline__ uint8_t png_byte(const PngGeom& g, const uint8_t* __restrict__ img, long long off) {
  if (off < 8) {
    const uint8_t sig[8] = {0x89, 0x50, 0x4E, 0x47, 0x0D, 0x0A, 0x1A, 0x0A};
    return sig[off];
  }
  if (off < 33) {
    const int o = (int)off - 8;
    if (o < 4) return o == 3 ? 13 : 0;
    if (o < 8) { const uint8_t t[4] = {'I', 'H', 'D', 'R'}; return t[o - 4]; }
    if (o < 12) return (uint8_t)((uint32_t)g.w >> (8 * (11 - o)));
    if (o < 16) return (uint8_t)((uint32_t)g.h >> (8 * (15 - o)));
    if (o == 16) return 8;
    if (o == 17) return 2;
    if (o < 21) return 0;
    return (uint8_t)(g.ihdr_crc >> (8 * (24 - o)));
  }
  const long long iend = PNG_IDAT + 12 + g.z;
  if (off >= iend) {
    const uint8_t e[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
    return e[off - iend];
  }
  const long long o = off - PNG_IDAT;
  if (o < 4) return (uint8_t)((unsigned long long)g.z >> (8 * (3 - o)));
  if (o < 8) { const uint8_t t[4] = {'I', 'D', 'A', 'T'}; return t[o - 4]; }
  long long s = o - 8;                                   // offset inside the zlib stream
  if (s == 0) return 0x78;
  if (s == 1) return 0x01;
  s -= 2;
  const long long body = g.raw + 5LL * g.nblk;
  if (s >= body) return 0;                               // adler32 / crc: filled by png_finish_kernel
  const long long k = s / (PNG_BLOCK + 5);
  const int b = (int)(s - k * (PNG_BLOCK + 5));
  if (b < 5) {
    const long long left = g.raw - k * PNG_BLOCK;
    const uint32_t len = (uint32_t)(left < PNG_BLOCK ? left : PNG_BLOCK);
    if (b == 0) return k == g.nblk - 1 ? 1 : 0;          // BFINAL, BTYPE = 00 (stored)
    if (b == 1) return (uint8_t)len;
    if (b == 2) return (uint8_t)(len >> 8);
    if (b == 3) return (uint8_t)~len;
    return (uint8_t)(~len >> 8);
  }
  const long long r = k * PNG_BLOCK + (b - 5);           // raw scanline stream index
  const long long line = 1 + 3LL * g.w;
  const long long y = r / line;
  const int c = (int)(r - y * line);
  if (c == 0) return 0;                                  // filter type 0 (None)
  return img[y * 3LL * g.w + (c - 1)];
}

// one thread per 4 file bytes; also one Adler-32 partial (sum, weighted sum) per scanline
__global__ void png_fill_kernel(const __grid_constant__ PngGeom g, const uint8_t* __restrict__ img, int nimg,
                                uint8_t* __restrict__ out, long long out_stride,
                                unsigned long long* __restrict__ adler_part) {
  const long long words = (g.total + 3) >> 2;
  const long long per_img = words + g.h;                 // then h scanline-checksum threads
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)nimg * per_img;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per_img);
    const long long t = i - (long long)n * per_img;
    const uint8_t* im = img + (long long)n * g.h * g.w * 3;
    if (t < words) {
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const long long off = t * 4 + b;
        if (off < g.total) v |= (uint32_t)png_byte(g, im, off) << (8 * b);
      }
      *reinterpret_cast<uint32_t*>(out + (long long)n * out_stride + t * 4) = v;
    } else {
      // Adler-32 partial of scanline y: A = sum d_i, B = sum (L - i) d_i over its L = 1 + 3w bytes
      // (the leading filter byte is 0: it only shifts the weights)
      const int y = (int)(t - words);
      const uint8_t* row = im + (long long)y * 3 * g.w;
      const int L = 1 + 3 * g.w;
      unsigned long long A = 0, B = 0;
      for (int c = 0; c < 3 * g.w; ++c) {
        const unsigned d = row[c];
        A += d;
        B += (unsigned long long)(L - 1 - c) * d;
      }
      adler_part[((long long)n * g.h + y) * 2] = A;
      adler_part[((long long)n * g.h + y) * 2 + 1] = B;
    }
  }
}

// one CTA per image: fold the scanline partials into the Adler-32, store it, CRC-32 the IDAT chunk
// (type + data) segment-parallel, fold the segment CRCs with the zero-shift operators, store the CRC
__global__ void __launch_bounds__(256)
png_finish_kernel(const __grid_constant__ PngGeom g, uint8_t* __restrict__ out, long long out_stride,
                  const unsigned long long* __restrict__ adler_part) {
  extern __shared__ uint32_t seg_crc[];
  uint8_t* f = out + (long long)blockIdx.x * out_stride;
  const long long adler_off = PNG_IDAT + 8 + 2 + g.raw + 5LL * g.nblk;
  if (threadIdx.x == 0) {
    const unsigned long long L = 1 + 3ULL * g.w;
    unsigned long long a = 1, b = 0;
    for (int y = 0; y < g.h; ++y) {
      const unsigned long long A = adler_part[((long long)blockIdx.x * g.h + y) * 2];
      const unsigned long long B = adler_part[((long long)blockIdx.x * g.h + y) * 2 + 1];
      b = (b + (L % 65521) * a + B) % 65521;
      a = (a + A) % 65521;
    }
    const uint32_t ad = (uint32_t)((b << 16) | a);
    f[adler_off + 0] = (uint8_t)(ad >> 24);
    f[adler_off + 1] = (uint8_t)(ad >> 16);
    f[adler_off + 2] = (uint8_t)(ad >> 8);
    f[adler_off + 3] = (uint8_t)ad;
  }
  __syncthreads();
  const long long n = 4 + g.z;                              // bytes under the CRC: "IDAT" + data
  const uint8_t* p = f + PNG_IDAT + 4;
  const int nseg = (int)((n + PNG_SEG - 1) / PNG_SEG);
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    const long long lo = (long long)s * PNG_SEG;
    const long long hi = lo + PNG_SEG < n ? lo + PNG_SEG : n;
    uint32_t c = 0xFFFFFFFFu;
    for (long long i = lo; i < hi; ++i) c = c_crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    seg_crc[s] = c ^ 0xFFFFFFFFu;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t crc = seg_crc[0];
    for (int s = 1; s < nseg; ++s)
      crc = dev_gf2_times(s == nseg - 1 ? g.op_tail : g.op_seg, crc) ^ seg_crc[s];
    const long long crc_off = PNG_IDAT + 8 + g.z;
    f[crc_off + 0] = (uint8_t)(crc >> 24);
    f[crc_off + 1] = (uint8_t)(crc >> 16);
    f[crc_off + 2] = (uint8_t)(crc >> 8);
    f[crc_off + 3] = (uint8_t)crc;
  }
}

}  // namespace dl

extern "C" long long dl_png_stored_size(int h, int w) {
  if (h <= 0 || w <= 0) return 0;
  return dl::png_geom(h, w).total;
}

extern "C" long long dl_png_stored_workspace_bytes(int nimg, int h) {
  return (long long)nimg * h * 2 * (long long)sizeof(unsigned long long);
}

extern "C" int dl_png_stored(const void* img_u8, int nimg, int h, int w, void* out, long long out_stride,
                             void* workspace, void* stream_) {
  using namespace dl;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  DL_CHECK_ARG(img_u8 && out && workspace && nimg > 0 && h > 0 && w > 0, "png_stored: bad args");
  PngGeom g = png_geom(h, w);
  DL_CHECK_ARG(out_stride >= ((g.total + 3) & ~3LL) && out_stride % 4 == 0,
               "png_stored: out_stride %lld must be a multiple of 4 and >= %lld", out_stride, (g.total + 3) & ~3LL);
  DL_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 3) == 0, "png_stored: out must be 4-byte aligned");
  DL_CHECK_ARG(g.z < (1LL << 31), "png_stored: image too large for one IDAT chunk");
  host_crc_init();
  static bool table_up[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!table_up[dev & 63]) {
    cudaError_t e = cudaMemcpyToSymbol(c_crc_table, h_crc_table, sizeof(h_crc_table));
    if (e != cudaSuccess) { set_error("png_stored: crc table upload: %s", cudaGetErrorString(e)); return 1; }
    table_up[dev & 63] = true;
  }
  {
    uint8_t ihdr[17] = {'I', 'H', 'D', 'R', 0, 0, 0, 0, 0, 0, 0, 0, 8, 2, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
      ihdr[4 + i] = (uint8_t)((uint32_t)w >> (8 * (3 - i)));
      ihdr[8 + i] = (uint8_t)((uint32_t)h >> (8 * (3 - i)));
    }
    g.ihdr_crc = host_crc(ihdr, 17);
  }
  const long long n = 4 + g.z;
  const int nseg = (int)((n + PNG_SEG - 1) / PNG_SEG);
  crc_zero_operator(g.op_seg, PNG_SEG);
  crc_zero_operator(g.op_tail, (unsigned long long)(n - (long long)(nseg - 1) * PNG_SEG));
  DL_CHECK_ARG(nseg * 4 <= 200 * 1024, "png_stored: image too large (CRC segments)");
  const long long threads = (long long)nimg * (((g.total + 3) >> 2) + h);
  long long blocks = (threads + 255) / 256;
  const long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  png_fill_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g, reinterpret_cast<const uint8_t*>(img_u8), nimg,
                                                        reinterpret_cast<uint8_t*>(out), out_stride,
                                                        reinterpret_cast<unsigned long long*>(workspace));
  if (check_launch("png_fill")) return 1;
  if (nseg * 4 > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(png_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nseg * 4);
    if (e != cudaSuccess) { set_error("png_stored: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return 1; }
  }
  png_finish_kernel<<<nimg, 256, nseg * 4, stream>>>(g, reinterpret_cast<uint8_t*>(out), out_stride,
                                                     reinterpret_cast<const unsigned long long*>(workspace));
  return check_launch("png_finish");
}
