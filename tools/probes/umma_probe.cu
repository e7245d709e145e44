// Probe: validate the hand-built tcgen05 shared-memory / instruction descriptors on a real B200 and
// measure raw tensor throughput of (a) tcgen05.mma issue loops and (b) legacy mma.sync, to choose
// the fast-path kernel designs.  Usage: umma_probe <test-id>   (one test per process: a bad descriptor
// traps the context).
//   0..7  : correctness, D[128xN] = A[128xK] * B[NxK]^T with the canonical no-swizzle layout
//           bit0: swap LBO/SBO meaning, bit1: B is MN-major (contract over rows), bit2: A is MN-major
//   10    : tcgen05 throughput (M=128,N=256,K=16 per instruction, long accumulate chain)
//   11    : mma.sync m16n8k16 bf16 throughput
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../cat-seg_b200/csrc/umma.cuh"

using namespace catseg::umma;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int iters) {
  for (int i = 0; i < iters; ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// A: logical [M=128][K] row-major floats (bf16-representable); B: logical [N][K] row-major.
// Storage in smem: canonical tiles; for MN-major variants the SAME bytes are described with swapped roles:
//   K-major  operand X[R][K]:  byte(r, k) = (k/8)*R*16 + r*16 + (k%8)*2          (LBO = R*16, SBO = 128)
//   MN-major operand X[R][K]:  byte(r, k) = (r/8)*K*16 + k*16 + (r%8)*2          (SBO = K*16, LBO = 128)
__global__ void gemm_probe(const float* A, const float* B, float* D, int N, int K, int variant, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int M = 128;
  const bool swap = variant & 1, b_mn = variant & 2, a_mn = variant & 4;
  // variant bit 3: the A tile starts at a row offset that is NOT a multiple of 8 rows (start address = base + 3*16):
  // needed by the implicit-GEMM convolutions, whose taps are row-shifted views of one shared-memory image
  uint8_t* sA = smem + ((variant & 8) ? 48 : 0);
  uint8_t* sB = smem + M * K * 2 + 128;
  int t = threadIdx.x;
  for (int i = t; i < M * K; i += blockDim.x) {
    int r = i / K, k = i % K;
    uint32_t off = a_mn ? (uint32_t)((r / 8) * K * 16 + k * 16 + (r % 8) * 2)
                        : (uint32_t)((k / 8) * M * 16 + r * 16 + (k % 8) * 2);
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = __float2bfloat16(A[i]);
  }
  for (int i = t; i < N * K; i += blockDim.x) {
    int r = i / K, k = i % K;
    uint32_t off = b_mn ? (uint32_t)((r / 8) * K * 16 + k * 16 + (r % 8) * 2)
                        : (uint32_t)((k / 8) * N * 16 + r * 16 + (k % 8) * 2);
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(B[i]);
  }
  if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (t < 32) tmem_alloc<256>(&tmem_base);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  uint32_t tm = tmem_base;
  if (t == 0) {
    uint32_t idesc = make_idesc_bf16(M, N, a_mn ? 1 : 0, b_mn ? 1 : 0);
    for (int k0 = 0; k0 < K; k0 += 16) {
      // K-major: K-chunk stride = R*16 (the "leading" offset), 8-row-group stride = 128 (the "stride" offset)
      // MN-major: 8-row(K)-group stride = 128, MN-chunk stride = K*16
      uint32_t a_addr, a_kdir, a_mndir, b_addr, b_kdir, b_mndir;
      if (!a_mn) { a_addr = smem_u32(sA) + (k0 / 8) * M * 16; a_kdir = M * 16; a_mndir = 128; }
      else       { a_addr = smem_u32(sA) + k0 * 16;           a_kdir = 128;    a_mndir = K * 16; }
      if (!b_mn) { b_addr = smem_u32(sB) + (k0 / 8) * N * 16; b_kdir = N * 16; b_mndir = 128; }
      else       { b_addr = smem_u32(sB) + k0 * 16;           b_kdir = 128;    b_mndir = K * 16; }
      uint64_t da = swap ? make_smem_desc(a_addr, a_mndir, a_kdir) : make_smem_desc(a_addr, a_kdir, a_mndir);
      uint64_t db = swap ? make_smem_desc(b_addr, b_mndir, b_kdir) : make_smem_desc(b_addr, b_kdir, b_mndir);
      mma_bf16_ss(tm, da, db, idesc, k0 > 0);
    }
    mma_commit(&bar);
  }
  bool ok = mbar_wait_bounded(&bar, 0, 1 << 22);
  fence_after_sync();
  if (!ok) { if (t == 0) *status = 1; }
  else {
    int warp = t >> 5, lane = t & 31;
    if (warp < 4) {
      for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = v[j];
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (t < 32) tmem_dealloc<256>(tm);
}

__global__ void tc_throughput(float* out, int iters, long long* cycles, int N, int lbo_a, int lbo_b, int commit_every = 0, int a_shift = 0, int b_shift = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t tmem_base;
  int t = threadIdx.x;
  for (int i = t; i < 20000; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (t == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); }
  if (t < 32) tmem_alloc<512>(&tmem_base);
  fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync();
  uint32_t tm = tmem_base;
  long long t0 = clock64();
  if (t == 0) {
    uint32_t idesc = make_idesc_bf16(128, N);
    uint32_t a0 = smem_u32(smem) + a_shift, b0 = smem_u32(smem + 40000) + b_shift;
    uint64_t da0 = make_smem_desc(a0, lbo_a, 128), db0 = make_smem_desc(b0, lbo_b, 128);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_bf16_ss(tm + (it & 1) * 256, da0 + (uint64_t)(k * 2 * (lbo_a >> 4)), db0 + (uint64_t)(k * 2 * (lbo_b >> 4)), idesc, 1);
      if (commit_every && (it % commit_every) == commit_every - 1) mma_commit(&bar2);   // nobody waits on bar2
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  long long t1 = clock64();
  fence_after_sync();
  if (t == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  float v[16];
  if (t < 128) { tmem_ld16(tm + ((uint32_t)((t >> 5) * 32) << 16), v); if (v[0] == 12345.f) out[t] = v[0]; }
  fence_before_sync(); __syncthreads();
  if (t < 32) tmem_dealloc<512>(tm);
}

// Commit probe (test 27): what does a tcgen05.commit in the MMA stream cost?  Thread 0 issues groups of `group` MMAs
// (M128, N, K16) and commits after every group according to `mode`:
//   0 no commit at all | 1 commit to ONE mbarrier nobody waits on | 2 commits round-robin over 8 mbarriers |
//   3 one mbarrier, warp 1 waits for every phase | 4 a single commit BEFORE the first MMA, none afterwards
__global__ void commit_probe(float* out, int ngroups, long long* cycles, int N, int group, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_base;
  int t = threadIdx.x;
  for (int i = t; i < 20000; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (t == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
  if (t < 32) tmem_alloc<512>(&tmem_base);
  fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync();
  uint32_t tm = tmem_base;
  const int lbo_a = 2064, lbo_b = N * 16;
  long long t0 = clock64();
  if (t == 0) {
    uint32_t idesc = make_idesc_bf16(128, N);
    uint64_t da0 = make_smem_desc(smem_u32(smem), lbo_a, 128), db0 = make_smem_desc(smem_u32(smem + 40000), lbo_b, 128);
    if (mode == 4) mma_commit(&bars[0]);
    for (int g = 0; g < ngroups; ++g) {
      for (int i = 0; i < group; ++i)
        mma_bf16_ss(tm + (g & 1) * 256, da0 + (uint64_t)((i & 3) * 2 * (lbo_a >> 4)), db0 + (uint64_t)((i & 3) * 2 * (lbo_b >> 4)), idesc, 1);
      if (mode == 1 || mode == 3) mma_commit(&bars[0]);
      else if (mode == 2) mma_commit(&bars[g & 7]);
    }
    mma_commit(&bar);
  } else if (t == 32 && mode == 3) {
    for (int g = 0; g < ngroups; ++g) (void)mbar_wait_bounded(&bars[0], (uint32_t)(g & 1), 1 << 12);   // bounded: a missed phase must not hang
  }
  mbar_wait(&bar, 0);
  long long t1 = clock64();
  fence_after_sync();
  if (t == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  float v[16];
  if (t < 128) { tmem_ld16(tm + ((uint32_t)((t >> 5) * 32) << 16), v); if (v[0] == 12345.f) out[t] = v[0]; }
  fence_before_sync(); __syncthreads();
  if (t < 32) tmem_dealloc<512>(tm);
}

// Contention probe: thread 0 issues a long MMA chain (M128 N128) while the other warps generate (mode bit 0) TMEM loads
// from other columns, (bit 1) shared-memory store+load traffic, (bit 2) global loads.  Reports cycles per MMA.
__global__ void contention(float* out, const float* gsrc, int nmma, long long* cycles, int mode, int random_data = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  __shared__ volatile int done;
  int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < 30000; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    // random bf16 pairs in [-2, 2): sign/mantissa random, exponent in {0x3e..0x3f}
    uint32_t rnd = (h & 0x807f807fu) | 0x3f003f00u | ((h >> 3) & 0x00800080u);
    reinterpret_cast<uint32_t*>(smem)[i] = random_data ? rnd : 0x3c003c00u;
  }
  if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); done = 0; }
  if (t < 32) tmem_alloc<512>(&tmem_base);
  fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync();
  uint32_t tm = tmem_base;
  const uint32_t idesc = make_idesc_bf16(128, 128);
  const uint64_t da0 = make_smem_desc(smem_u32(smem), 2064, 128), db0 = make_smem_desc(smem_u32(smem + 40000), 2048, 128);
  long long t0 = clock64();
  float acc = 0.f;
  if (warp == 0) {
    if (elect_one()) {
      for (int g = 0; g < nmma / 8; ++g) {
#pragma unroll
        for (int k = 0; k < 8; ++k) mma_bf16_ss(tm + (g & 1) * 128, da0 + (uint64_t)(k * 2 * 129), db0 + (uint64_t)(k * 2 * 128), idesc, 1);
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (t == 0) done = 1;
  } else {
    const uint32_t lane_addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    float* sp = reinterpret_cast<float*>(smem + 120000) + t * 4;
    while (!done) {
      if (mode & 1) { float v[32]; tmem_ld32(lane_addr + ((warp >> 2) & 3) * 32, v); acc += v[0] + v[31]; }
      if (mode & 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { *reinterpret_cast<float4*>(sp + i * 2048) = make_float4(acc, acc, acc, acc); acc += sp[i * 2048 + 1]; }
      }
      if (mode & 4) { acc += __ldg(gsrc + ((t * 32 + (int)acc) & 0xfffff)); }
    }
  }
  long long t1 = clock64();
  fence_after_sync();
  if (t == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 12345.f) out[t] = acc;
  fence_before_sync(); __syncthreads();
  if (t < 32) tmem_dealloc<512>(tm);
}

// Issue-cost probe: GEMMs of 8 MMAs (N=128) with k-dependent descriptors + a commit per GEMM, issued either by
// `if (threadIdx.x == 0)` (mode 0) or by warp 0 under elect.sync in a warp-uniform branch (mode 1).
__global__ void issue_cost(float* out, int ngemm, long long* cycles, int mode, int a_off = 0, int b_off = 40000) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base;
  int t = threadIdx.x;
  for (int i = t; i < 55000; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (t == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); }
  if (t < 32) tmem_alloc<512>(&tmem_base);
  fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync();
  uint32_t tm = tmem_base;
  const uint32_t idesc = make_idesc_bf16(128, 128);
  const uint64_t da0 = make_smem_desc(smem_u32(smem + a_off), 2064, 128), db0 = make_smem_desc(smem_u32(smem + b_off), 2048, 128);
  const int warp_u = __shfl_sync(0xffffffffu, t >> 5, 0);
  long long t0 = clock64();
  if (mode == 0) {
    if (t == 0) {
      for (int g = 0; g < ngemm; ++g) {
        uint64_t wd = db0 + (uint64_t)((uint32_t)(g % 3) * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) mma_bf16_ss(tm + (g & 3) * 128, da0 + (uint64_t)(k * 2 * 129), wd + (uint64_t)(k * 2 * 128), idesc, k > 0);
        mma_commit(&bar2);
      }
      mma_commit(&bar);
    }
  } else {
    if (warp_u == 0) {
      for (int g = 0; g < ngemm; ++g) {
        uint64_t wd = db0 + (uint64_t)((uint32_t)(g % 3) * 8);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) mma_bf16_ss(tm + (g & 3) * 128, da0 + (uint64_t)(k * 2 * 129), wd + (uint64_t)(k * 2 * 128), idesc, k > 0);
          mma_commit(&bar2);
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(&bar);
      __syncwarp();
    }
  }
  long long t_issue = clock64();
  mbar_wait(&bar, 0);
  long long t1 = clock64();
  fence_after_sync();
  if (t == 0 && blockIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t_issue - t0; }
  float v[16];
  if (t < 128) { tmem_ld16(tm + ((uint32_t)((t >> 5) * 32) << 16), v); if (v[0] == 12345.f) out[t] = v[0]; }
  fence_before_sync(); __syncthreads();
  if (t < 32) tmem_dealloc<512>(tm);
}


// Commit-latency probe: thread 0 issues NG GEMMs (8 MMAs, N=128), each with its own commit barrier, then every thread
// polls the barriers in order; thread 0 records when each flip becomes visible.  mode 0: other warps only poll;
// mode 1: the other warps first run an ALU-bound loop (like an activation epilogue); mode 2: they stream STS.128.
__global__ void commit_latency(float* out, long long* cycles, int mode, int ng) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tmem_base;
  int t = threadIdx.x;
  for (int i = t; i < 40000; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (t == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
  if (t < 32) tmem_alloc<512>(&tmem_base);
  fence_proxy_async(); fence_before_sync(); __syncthreads(); fence_after_sync();
  uint32_t tm = tmem_base;
  const uint32_t idesc = make_idesc_bf16(128, 128);
  const uint64_t da0 = make_smem_desc(smem_u32(smem), 2064, 128), db0 = make_smem_desc(smem_u32(smem + 40000), 2048, 128);
  const int warp_u = __shfl_sync(0xffffffffu, t >> 5, 0);
  long long t0 = clock64(), tissue = 0;
  if (warp_u == 0) {
    if (elect_one()) {
      for (int g = 0; g < ng; ++g) {
#pragma unroll
        for (int k = 0; k < 8; ++k) mma_bf16_ss(tm + (g & 3) * 128, da0 + (uint64_t)(k * 2 * 129), db0 + (uint64_t)(k * 2 * 128), idesc, k > 0);
        mma_commit(&bars[g]);
      }
    }
    __syncwarp();
    tissue = clock64();
  } else if (mode == 1) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = (float)(t + i);
    for (int it = 0; it < 400; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 1.0001f, 0.5f);
    float sacc = 0.f; for (int i = 0; i < 8; ++i) sacc += a[i];
    if (sacc == 12345.f) out[t] = sacc;
  } else if (mode == 2) {
    uint4* dst = reinterpret_cast<uint4*>(smem + 90000);
    for (int it = 0; it < 200; ++it) dst[(t + it * 32) & 4095] = make_uint4(t, it, 0, 0);
  }
  long long tb[8];
  for (int g = 0; g < ng; ++g) { mbar_wait(&bars[g], 0); tb[g] = clock64(); }
  fence_after_sync();
  if (t == 0 && blockIdx.x == 0) cycles[0] = tissue - t0;
  if (t == 32 && blockIdx.x == 0) for (int g = 0; g < ng; ++g) cycles[1 + g] = tb[g] - t0;   // observed by a warp that does not issue
  float v[16];
  if (t < 128) { tmem_ld16(tm + ((uint32_t)((t >> 5) * 32) << 16), v); if (v[0] == 12345.f) out[t] = v[0]; }
  fence_before_sync(); __syncthreads();
  if (t < 32) tmem_dealloc<512>(tm);
}

__global__ void mmasync_throughput(float* out, int iters, long long* cycles) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a0 = 0x3c003c00u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3c003c00u, b1 = b0 + 5;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 12345.f) out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
  int id = argc > 1 ? atoi(argv[1]) : 0;
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  if (id < 16) {
    const int M = 128, K = 64;
    const int Ns[3] = {64, 144, 256};
    for (int ni = 0; ni < 3; ++ni) {
      int N = Ns[ni];
      std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N);
      srand(1 + ni);
      for (auto& v : A) v = bf16r((rand() % 2001 - 1000) / 1000.f);
      for (auto& v : B) v = bf16r((rand() % 2001 - 1000) / 1000.f);
      for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; R[m * N + n] = (float)s; }
      float *dA, *dB, *dD; int* dS;
      CHECK(cudaMalloc(&dA, A.size() * 4)); CHECK(cudaMalloc(&dB, B.size() * 4)); CHECK(cudaMalloc(&dD, D.size() * 4)); CHECK(cudaMalloc(&dS, 4));
      CHECK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
      CHECK(cudaMemset(dD, 0, D.size() * 4)); CHECK(cudaMemset(dS, 0, 4));
      size_t smem = (size_t)(M + N) * K * 2 + 256;
      CHECK(cudaFuncSetAttribute(gemm_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      gemm_probe<<<1, 128, smem>>>(dA, dB, dD, N, K, id, dS);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("PROBE id=%d N=%d CUDA-ERROR %s\n", id, N, cudaGetErrorString(e)); return 3; }
      int st; CHECK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost)); CHECK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0; for (size_t i = 0; i < D.size(); ++i) maxerr = fmax(maxerr, fabs((double)D[i] - R[i]));
      printf("PROBE id=%d swap=%d b_mn=%d a_mn=%d N=%d timeout=%d maxerr=%.5f %s\n", id, id & 1, (id >> 1) & 1, (id >> 2) & 1, N, st, maxerr,
             (!st && maxerr < 1e-3) ? "OK" : "BAD");
    }
  } else {
    float* out; long long* cyc; CHECK(cudaMalloc(&out, 4096)); CHECK(cudaMalloc(&cyc, 8));
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (id == 25) {
      size_t smem = 160000;
      CHECK(cudaFuncSetAttribute(commit_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      long long* cyc2; CHECK(cudaMalloc(&cyc2, 128));
      const char* names[] = {"poll only", "ALU loop", "STS stream"};
      for (int ng = 2; ng <= 4; ng += 2)
        for (int mode = 0; mode < 3; ++mode)
          for (int nt = 128; nt <= 512; nt += 384) {
            commit_latency<<<prop.multiProcessorCount, nt, smem>>>(out, cyc2, mode, ng); CHECK(cudaDeviceSynchronize());
            commit_latency<<<prop.multiProcessorCount, nt, smem>>>(out, cyc2, mode, ng); CHECK(cudaDeviceSynchronize());
            long long c[9]; CHECK(cudaMemcpy(c, cyc2, 72, cudaMemcpyDeviceToHost));
            printf("PROBE commit latency, %d GEMMs x 8 MMAs, %3d threads, others %-10s: issue done %lld |", ng, nt, names[mode], c[0]);
            for (int g = 0; g < ng; ++g) printf(" bar%d %lld", g, c[1 + g]);
            printf(" cycles\n");
          }
    } else if (id == 24) {
      size_t smem = 200000;
      CHECK(cudaFuncSetAttribute(contention, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float* gsrc; CHECK(cudaMalloc(&gsrc, 4 << 20)); CHECK(cudaMemset(gsrc, 0, 4 << 20));
      long long* cyc2; CHECK(cudaMalloc(&cyc2, 16));
      const char* names[] = {"idle warps", "TMEM loads", "SMEM st/ld", "TMEM+SMEM", "global loads", "TMEM+global", "SMEM+global", "all"};
      for (int mode = 0; mode < 8; ++mode) {
        contention<<<prop.multiProcessorCount, 512, smem>>>(out, gsrc, 16000, cyc2, mode); CHECK(cudaDeviceSynchronize());
        long long c; CHECK(cudaMemcpy(&c, cyc2, 8, cudaMemcpyDeviceToHost));
        printf("PROBE contention, 15 other warps doing %-14s: %.1f cycles per MMA (M128 N128 K16)\n", names[mode], (double)c / 16000.0);
      }
      for (int rep = 0; rep < 2; ++rep) {
        contention<<<prop.multiProcessorCount, 512, smem>>>(out, gsrc, 64000, cyc2, 0, 1); CHECK(cudaDeviceSynchronize());
        long long c; CHECK(cudaMemcpy(&c, cyc2, 8, cudaMemcpyDeviceToHost));
        printf("PROBE random bf16 operands (idle warps): %.1f cycles per MMA (M128 N128 K16)\n", (double)c / 64000.0);
      }
    } else if (id == 23) {
      size_t smem = 100000;
      CHECK(cudaFuncSetAttribute(issue_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      long long* cyc2; CHECK(cudaMalloc(&cyc2, 16));
      smem = 225000;
      CHECK(cudaFuncSetAttribute(issue_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int offs[][2] = {{0, 40000}, {100000, 40000}, {135168, 40000}, {170000, 40000}, {0, 135168}, {135168, 175104}, {98304, 0}, {131328, 0}, {164352, 32768}};
      for (auto& o : offs) {
        issue_cost<<<prop.multiProcessorCount, 512, smem>>>(out, 2000, cyc2, 0, o[0], o[1]); CHECK(cudaDeviceSynchronize());
        long long c[2]; CHECK(cudaMemcpy(c, cyc2, 16, cudaMemcpyDeviceToHost));
        printf("PROBE operand placement A@%d B@%d: %.1f cycles per MMA (M128 N128)\n", o[0], o[1], (double)c[0] / (2000.0 * 8));
      }
      for (int mode = 0; mode < 2; ++mode) {
        issue_cost<<<prop.multiProcessorCount, 512, smem>>>(out, 2000, cyc2, mode); CHECK(cudaDeviceSynchronize());
        long long c[2]; CHECK(cudaMemcpy(c, cyc2, 16, cudaMemcpyDeviceToHost));
        printf("PROBE issue cost (%s): %.1f cycles per MMA end-to-end, %.1f cycles per MMA spent issuing (8-MMA GEMMs + commit)\n",
               mode == 0 ? "if (tid == 0)" : "warp 0 + elect.sync", (double)c[0] / (2000.0 * 8), (double)c[1] / (2000.0 * 8));
      }
    } else if (id == 27) {
      size_t smem = 100000;
      CHECK(cudaFuncSetAttribute(commit_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int total = 8192;                       // MMAs per run
      const int Ns[] = {128, 64, 256};
      const int groups[] = {4, 8, 16, 32, 64, 256, 1024};
      for (int N : Ns)
        for (int mode = 0; mode < 5; ++mode)
          for (int group : groups) {
            if ((mode == 0 || mode == 4) && group != 16) continue;
            if (N != 128 && (mode == 2 || mode == 3 || (group != 8 && group != 32 && group != 256))) { if (!(mode == 0 || mode == 4)) continue; }
            commit_probe<<<prop.multiProcessorCount, 128, smem>>>(out, total / group, cyc, N, group, mode); CHECK(cudaDeviceSynchronize());
            long long c; CHECK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
            printf("PROBE commit N=%d mode=%d (%s) group=%d MMAs: %.1f cycles per MMA\n", N, mode,
                   mode == 0 ? "no commit" : mode == 1 ? "one barrier, no waiter" : mode == 2 ? "8 barriers round-robin" : mode == 3 ? "one barrier, a waiting warp" : "single commit up front",
                   group, (double)c / total);
          }
    } else if (id == 26) {
      // row-shifted operand starts (implicit-GEMM convolution taps): does an A / B tile whose first row is not a multiple
      // of 8 rows (start address not 128-byte aligned) cost extra shared-memory wavefronts?
      size_t smem = 100000;
      CHECK(cudaFuncSetAttribute(tc_throughput, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int cfgs[][5] = {{64, 2064, 1024, 0, 0}, {64, 2064, 1024, 16, 0}, {64, 2064, 1024, 64, 0}, {64, 2064, 1024, 128, 0}, {64, 2064, 1024, 816, 0},
                             {64, 2048, 1024, 0, 0}, {64, 2048, 1024, 16, 0}, {64, 6416, 1024, 816, 0}, {64, 6416, 1024, 800, 0},
                             {32, 2064, 512, 0, 0}, {32, 2064, 512, 16, 0}, {128, 2064, 2048, 0, 0}, {128, 2064, 2048, 16, 0}, {256, 2064, 4096, 0, 0}, {256, 2064, 4096, 16, 0},
                             {64, 2064, 1024, 0, 16}, {144, 2048, 2320, 0, 16},
                             // decoder D1 / D2 shapes: the lo MMA (N = 64) reads the first 64 rows of a 128-row [Wh | Wl] image (lboB = 2048)
                             {64, 2064, 2048, 0, 0}, {64, 4240, 2048, 432, 0}, {128, 4240, 2048, 432, 0}, {64, 8080, 2048, 816, 0}, {128, 8080, 2048, 816, 0},
                             {32, 8080, 1024, 816, 0}, {64, 8080, 1024, 816, 0}};
      for (auto& c5 : cfgs) {
        tc_throughput<<<prop.multiProcessorCount, 128, smem>>>(out, iters, cyc, c5[0], c5[1], c5[2], 0, c5[3], c5[4]); CHECK(cudaDeviceSynchronize());
        long long c; CHECK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        printf("PROBE tcgen05 M128 N%d lboA=%d lboB=%d A-start +%d B, B-start +%d B: %.1f cycles per MMA(K=16)\n", c5[0], c5[1], c5[2], c5[3], c5[4], (double)c / (4.0 * iters));
      }
    } else if (id == 20) {
      size_t smem = 100000;
      CHECK(cudaFuncSetAttribute(tc_throughput, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // (N, lbo_a, lbo_b): dense vs padded (not 128-byte aligned) K-chunk strides
      const int cfgs[][3] = {{256, 2048, 4096}, {128, 2048, 2048}, {128, 2064, 2048}, {128, 2064, 2064}, {144, 2048, 2320}, {144, 2048, 2304},
                             {64, 2048, 1024}, {32, 2048, 512}, {32, 2320, 512}, {16, 2048, 256}};
      int nthreads = argc > 2 ? atoi(argv[2]) : 128;     // all threads but one spin on the mbarrier while the MMAs run
      printf("PROBE threads per CTA = %d\n", nthreads);
      for (auto& c3 : cfgs) {
        tc_throughput<<<prop.multiProcessorCount, nthreads, smem>>>(out, 100, cyc, c3[0], c3[1], c3[2]); CHECK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        tc_throughput<<<prop.multiProcessorCount, nthreads, smem>>>(out, iters, cyc, c3[0], c3[1], c3[2]);
        cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); long long c; CHECK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        double flops = 2.0 * 128 * c3[0] * 16 * 4 * iters * prop.multiProcessorCount;
        printf("PROBE tcgen05 M128 N%d lboA=%d lboB=%d: %.1f TFLOP/s, %.1f cycles per MMA(K=16)\n", c3[0], c3[1], c3[2], flops / ms * 1e-9, (double)c / (4.0 * iters));
      }
      for (int ce = 1; ce <= 4; ce *= 2) {      // a tcgen05.commit after every 4*ce MMAs (as the fused kernels do per GEMM)
        tc_throughput<<<prop.multiProcessorCount, nthreads, smem>>>(out, iters, cyc, 128, 2064, 2048, ce); CHECK(cudaDeviceSynchronize());
        long long c; CHECK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        printf("PROBE tcgen05 M128 N128 with a commit every %d MMAs: %.1f cycles per MMA(K=16)\n", 4 * ce, (double)c / (4.0 * iters));
      }
    } else {
      for (int warps = 4; warps <= 16; warps *= 2) {
        mmasync_throughput<<<prop.multiProcessorCount, warps * 32>>>(out, 100, cyc); CHECK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        mmasync_throughput<<<prop.multiProcessorCount, warps * 32>>>(out, iters, cyc);
        cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 16 * 8 * 16 * 8.0 * iters * warps * prop.multiProcessorCount;
        printf("PROBE mma.sync m16n8k16 bf16, %d warps/SM: %.3f ms, %.1f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
      }
    }
  }
  return 0;
}
