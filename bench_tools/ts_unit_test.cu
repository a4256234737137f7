// Numerics harness for the "stacked rows" 3x3 convolution on tcgen05 with the weights in tensor memory:
//   D[128 = (output row y | output row y+1) x 64 channels, 64 pixels] += A[tmem: stacked weights 128 x 16] * B[smem: 64 pixels x 16 ch]
// For input row offset r = 0..3 and column offset kx = 0..2 the upper half of A holds W(r, kx) (output row y
// sees input row y+r through tap ky = r) and the lower half W(r-1, kx) (output row y+1 sees the same input row
// through tap ky = r-1); halves whose tap does not exist are zero.  12 such groups x 4 slices of 16 input
// channels = 48 MMAs for 2 x 64 output pixels x 64 channels.  One CTA computes one such unit and the host
// checks it against a plain loop.  What this pins down before the trunk kernel is rewritten around it:
//   - the tensor-memory layout of an fp16 A operand (written with tcgen05.st.32x32b.x8: lane = row, 8 columns = 16 K values)
//   - B windows over a 4 x 66 pixel halo (SWIZZLE_128B rows of 128 B, start address shifted by whole pixels)
//   - tcgen05.ld.16x256b register layout, stmatrix.trans / ldmatrix.trans for the channel-major -> NHWC transposition
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I joshupscale_b200/csrc/kernels \
//        bench_tools/ts_unit_test.cu -o bench_tools/_ts_unit_test
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "tc_common.cuh"

using namespace ju::tc;

constexpr int kN = 64;              // pixels per output row of the unit
constexpr int kPitch = kN + 2;      // halo pixels per input row
constexpr int kHaloPix = 4 * kPitch;
constexpr uint32_t kAccCol = 0;
constexpr uint32_t kWCol = 128;     // weights: (group * 4 + slice) * 8 columns

__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
	    "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
	    : "memory");
}
// same, with a lane mask: bit i of word w set = lane 32 * w + i of D is NOT written
__device__ __forceinline__ void umma_ts_f16_masked(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc,
    uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(d_tmem),
	    "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
	    : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
	    "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
	    : "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
	      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
	      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
	      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
	    : "r"(taddr)
	    : "memory");
}
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
	asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
	             : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
	asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
// byte offset of (pixel line `line`, channel `ch`) in a SWIZZLE_128B tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t swz(uint32_t line, uint32_t ch) { return line * 128u + ((((ch >> 3) ^ (line & 7u)) << 4) | ((ch & 7u) << 1)); }

// x: [4][66][64] fp16 halo, w: [9][64 cout][64 cin] fp16, bias [64], res: [2][64][64] fp16, out: [2][64][64] fp16
// probe: [128 threads][32] what tcgen05.ld.16x256b.x8 returned for a TMEM pattern lane * 1000 + column
__global__ void __launch_bounds__(128, 1)
unit_kernel(const __half *x, const __half *w, const float *bias, const __half *res, __half *out, float *probe) {
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint32_t tmem_slot;
	__shared__ __align__(8) unsigned long long bar;
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	unsigned char *gen = smem_raw + (base - smem_u32(smem_raw));
	const uint32_t halo = 0;                       // 4 * 66 * 128 = 33792 B
	const uint32_t rtile = 34 * 1024;              // residual tile, 2 * 64 * 128 = 16 KB
	const uint32_t stile = rtile + 16 * 1024;      // output staging tile
	const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (threadIdx.x == 0) {
		mbar_init(smem_u32(&bar), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int i = threadIdx.x; i < kHaloPix * 64; i += 128) {
		const int p = i / 64, ch = i % 64;
		*reinterpret_cast<__half *>(gen + halo + swz(p, ch)) = x[i];
	}
	for (int i = threadIdx.x; i < 2 * kN * 64; i += 128) {
		const int p = i / 64, ch = i % 64;
		*reinterpret_cast<__half *>(gen + rtile + swz(p, ch)) = res[i];
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();
	const uint32_t tmem = tmem_slot;
	const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;

	// ---- probe of the 16x256b load layout -------------------------------------------------
	{
		for (int cb = 0; cb < 8; ++cb) {
			uint32_t v[8];
			for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(static_cast<float>(threadIdx.x * 1000 + cb * 8 + e));
			tmem_st8(tmem + lane_base + kAccCol + cb * 8, v);
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
		__syncwarp();
		uint32_t r[32];
		tmem_ld_16x256b_x8(tmem + lane_base + kAccCol, r);
		tmem_ld_wait();
		for (int e = 0; e < 32; ++e) probe[threadIdx.x * 32 + e] = __uint_as_float(r[e]);
		__syncwarp();
	}

	// ---- weights into tensor memory: thread m owns row m of every stacked slice ------------
	// 9 stored groups: (kx, t = 0) = [W(0, kx); W(2, kx)] serves input row 0 (upper half only) AND input row 3
	// (lower half only) through the output-lane mask; (kx, t = 1, 2) = [W(t, kx); W(t - 1, kx)] serve input rows 1, 2
	{
		const int m = threadIdx.x, h = m / 64, c = m % 64;
		for (int g = 0; g < 9; ++g) {
			const int kx = g / 3, t = g % 3;
			const int ky = t == 0 ? (h ? 2 : 0) : t - h;
			for (int j = 0; j < 4; ++j) {
				uint32_t v[8];
				const uint4 *src = reinterpret_cast<const uint4 *>(w + ((ky * 3 + kx) * 64 + c) * 64 + j * 16);
				const uint4 a = src[0], b = src[1];
				v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
				tmem_st8(tmem + lane_base + kWCol + (g * 4 + j) * 8, v);
			}
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();

	// ---- 48 MMAs ---------------------------------------------------------------------------
	if (warp == 0) {
		const uint32_t idesc = make_idesc(kN);
		if (elect_one_sync()) {
			bool first = true;
			for (int kx = 0; kx < 3; ++kx) {
				for (int ri = 0; ri < 4; ++ri) {
					// an unmasked group goes first: the masked ones must not be the ones that initialise D
					const int r = ri == 0 ? 1 : (ri == 1 ? 2 : (ri == 2 ? 0 : 3));
					const int g = kx * 3 + (r == 0 || r == 3 ? 0 : r);
					for (int j = 0; j < 4; ++j) {
						const uint64_t bd = make_smem_desc(base + halo + (r * kPitch + kx) * 128 + j * 32, 1024u, 0);
						const uint32_t a = tmem + kWCol + (g * 4 + j) * 8;
						if (r == 0) umma_ts_f16_masked(tmem + kAccCol, a, bd, idesc, !first, 0u, 0u, 0xffffffffu, 0xffffffffu);
						else if (r == 3) umma_ts_f16_masked(tmem + kAccCol, a, bd, idesc, !first, 0xffffffffu, 0xffffffffu, 0u, 0u);
						else umma_ts_f16(tmem + kAccCol, a, bd, idesc, !first);
						first = false;
					}
				}
			}
			umma_commit(smem_u32(&bar));
		}
		__syncwarp();
	}
	while (!mbar_try_wait(smem_u32(&bar), 0)) {
	}
	tcgen05_fence_after();

	// ---- epilogue: warp q = (output row q / 2, channels 32 * (q % 2) ..) --------------------
	{
		const int orow = warp >> 1;
		for (int hb = 0; hb < 2; ++hb) {
			uint32_t r[32];
			tmem_ld_16x256b_x8(tmem + ((static_cast<uint32_t>(warp * 32 + hb * 16)) << 16) + kAccCol, r);
			tmem_ld_wait();
			const int c_lo = 32 * (warp & 1) + 16 * hb + lane / 4, c_hi = c_lo + 8;
			const float b_lo = bias[c_lo], b_hi = bias[c_hi];
			// pixel blocks in pairs: one stmatrix.x4 = {c_lo block, c_hi block} x {pixel block pb, pb + 1}
			for (int pb = 0; pb < 8; pb += 2) {
				// address of this lane's matrix row: matrix i = lane / 8 -> (channel block i & 1, pixel block pb + (i >> 1)), row = pixel
				const int mi = lane / 8, mr = lane % 8;
				const uint32_t line = static_cast<uint32_t>(orow * kN + (pb + (mi >> 1)) * 8 + mr);
				const uint32_t cbase = static_cast<uint32_t>(32 * (warp & 1) + 16 * hb + 8 * (mi & 1));
				const uint32_t off = swz(line, cbase);
				uint32_t q0, q1, q2, q3;
				ldmatrix_x4_trans(base + rtile + off, q0, q1, q2, q3);
				uint32_t o[4];
				const uint32_t qq[4] = {q0, q1, q2, q3};
				for (int i = 0; i < 4; ++i) {
					const int blk = pb + (i >> 1);
					const bool hi = i & 1;
					const float a0 = __uint_as_float(r[blk * 4 + (hi ? 2 : 0)]), a1 = __uint_as_float(r[blk * 4 + (hi ? 3 : 1)]);
					const float2 rr = __half22float2(*reinterpret_cast<const __half2 *>(&qq[i]));
					const float v0 = fmaxf(a0 + (hi ? b_hi : b_lo) + rr.x, 0.f), v1 = fmaxf(a1 + (hi ? b_hi : b_lo) + rr.y, 0.f);
					const __half2 hh = __floats2half2_rn(v0, v1);
					o[i] = *reinterpret_cast<const uint32_t *>(&hh);
				}
				stmatrix_x4_trans(base + stile + off, o[0], o[1], o[2], o[3]);
			}
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 2 * kN * 64; i += 128) {
		const int p = i / 64, ch = i % 64;
		out[i] = *reinterpret_cast<const __half *>(gen + stile + swz(p, ch));
	}
	tcgen05_fence_before();
	__syncthreads();
	if (warp == 0) {
		tcgen05_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
	}
}

int main() {
	std::vector<__half> x(kHaloPix * 64), w(9 * 64 * 64), res(2 * kN * 64), out(2 * kN * 64);
	std::vector<float> bias(64), probe(128 * 32);
	uint32_t s = 12345;
	auto rnd = [&]() {
		s = s * 1664525u + 1013904223u;
		return (static_cast<float>((s >> 8) & 0xffff) / 65536.f) - 0.5f;
	};
	for (auto &v : x) v = __float2half(rnd());
	for (auto &v : w) v = __float2half(rnd() * 0.2f);
	for (auto &v : res) v = __float2half(rnd());
	for (auto &v : bias) v = rnd();
	__half *dx, *dw, *dres, *dout;
	float *dbias, *dprobe;
	cudaMalloc(&dx, x.size() * 2);
	cudaMalloc(&dw, w.size() * 2);
	cudaMalloc(&dres, res.size() * 2);
	cudaMalloc(&dout, out.size() * 2);
	cudaMalloc(&dbias, 64 * 4);
	cudaMalloc(&dprobe, probe.size() * 4);
	cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice);
	cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
	cudaMemcpy(dres, res.data(), res.size() * 2, cudaMemcpyHostToDevice);
	cudaMemcpy(dbias, bias.data(), 64 * 4, cudaMemcpyHostToDevice);
	cudaMemset(dout, 0, out.size() * 2);
	const int smem = 1024 + 34 * 1024 + 32 * 1024;
	cudaFuncSetAttribute(unit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	unit_kernel<<<1, 128, smem>>>(dx, dw, dbias, dres, dout, dprobe);
	cudaError_t e = cudaDeviceSynchronize();
	printf("kernel: %s\n", cudaGetErrorString(e));
	cudaMemcpy(out.data(), dout, out.size() * 2, cudaMemcpyDeviceToHost);
	cudaMemcpy(probe.data(), dprobe, probe.size() * 4, cudaMemcpyDeviceToHost);
	// probe: value = lane * 1000 + column
	printf("16x256b.x8 layout (thread: reg -> lane,col):\n");
	for (int t : {0, 1, 4, 5, 31, 32, 33}) {
		printf("  t%-3d", t);
		for (int r = 0; r < 10; ++r) {
			const int v = static_cast<int>(probe[t * 32 + r]);
			printf(" r%d=(%d,%d)", r, v / 1000, v % 1000);
		}
		printf("\n");
	}
	int layout_bad = 0;
	for (int t = 0; t < 128; ++t) {
		for (int r = 0; r < 32; ++r) {
			const int v = static_cast<int>(probe[t * 32 + r]);
			const int want_lane = (t / 32) * 32 + (t % 32) / 4 + ((r & 2) ? 8 : 0), want_col = (r / 4) * 8 + 2 * (t % 4) + (r & 1);
			if (v / 1000 != want_lane || v % 1000 != want_col) ++layout_bad;
		}
	}
	printf("16x256b layout mismatches against the assumed mapping: %d\n", layout_bad);
	double max_err = 0;
	int bad = 0;
	for (int orow = 0; orow < 2; ++orow) {
		for (int px = 0; px < kN; ++px) {
			for (int co = 0; co < 64; ++co) {
				double acc = bias[co];
				for (int ky = 0; ky < 3; ++ky) {
					for (int kx = 0; kx < 3; ++kx) {
						for (int ci = 0; ci < 64; ++ci) {
							acc += static_cast<double>(__half2float(w[((ky * 3 + kx) * 64 + co) * 64 + ci])) *
							       __half2float(x[((orow + ky) * kPitch + px + kx) * 64 + ci]);
						}
					}
				}
				acc += __half2float(res[(orow * kN + px) * 64 + co]);
				if (acc < 0) acc = 0;
				const double got = __half2float(out[(orow * kN + px) * 64 + co]);
				const double err = std::fabs(got - acc);
				if (err > max_err) max_err = err;
				if (err > 0.02 + 0.004 * std::fabs(acc)) {
					if (bad < 8) printf("  mismatch row %d px %d ch %d: got %.4f want %.4f\n", orow, px, co, got, acc);
					++bad;
				}
			}
		}
	}
	printf("unit: max abs err %.5f, %d of %d outside tolerance -> %s\n", max_err, bad, 2 * kN * 64, bad == 0 && layout_bad == 0 ? "PASS" : "FAIL");
	return bad == 0 && layout_bad == 0 ? 0 : 1;
}
