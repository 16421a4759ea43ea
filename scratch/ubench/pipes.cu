// Throughput microbenchmark for the integer ops the lifting kernels are made of (B200, sm_100a).
// Each kernel runs N dependent-chain-free ops per thread on 8 independent accumulators; reports ops/clk/SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b)
{
	uint32_t d;
	if (OP == 0) asm volatile("mad.lo.s32 %0, %1, %2, %1;" : "=r"(d) : "r"(a), "r"(b));                 // IMAD
	if (OP == 1) asm volatile("mul.hi.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));                      // IMAD.HI
	if (OP == 2) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %1;" : "=r"(d) : "r"(a), "r"(b));             // IDP.2A
	if (OP == 3) asm volatile("dp4a.s32.s32 %0, %1, %2, %1;" : "=r"(d) : "r"(a), "r"(b));                // IDP.4A
	if (OP == 4) asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(d) : "r"(a), "r"(b));                // PRMT
	if (OP == 5) asm volatile("shr.s32 %0, %1, 3;\n\tadd.s32 %0, %0, %2;" : "=r"(d) : "r"(a), "r"(b));   // LEA.HI.SX32 ?
	if (OP == 6) asm volatile("shr.s32 %0, %1, 16;" : "=r"(d) : "r"(a));                                 // SHF
	if (OP == 7) asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(d) : "r"(a), "r"(b));              // LOP3
	if (OP == 8) d = __vadd2(a, b);                                                                          // VIADD.16x2
	if (OP == 9) asm volatile("add.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));                         // IADD3 / IMAD.IADD
	if (OP == 10) asm volatile("mad.hi.s32 %0, %1, %2, %1;" : "=r"(d) : "r"(a), "r"(b));                 // IMAD.HI with addend
	if (OP == 11) asm volatile("cvt.s32.s16 %0, %1;" : "=r"(d) : "h"((unsigned short)a));                // sign extend
	if (OP == 12) d = __vmaxs2(a, b);                                                                        // VIMNMX.S16x2
	if (OP == 13) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(*(unsigned long long*)&d) : "r"(a), "r"(b));
	if (OP == 14) d = __viaddmax_s16x2(a, b, 0x80008000u);                                                  // VIADDMNMX.S16x2 (DPX)
	if (OP == 15) d = __vimax3_s16x2(a, b, 0x80008000u);                                                    // VIMNMX3 16x2 (DPX)
	if (OP == 16) asm volatile("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(d) : "r"(a), "r"(b));               // HFMA2
	if (OP == 17) asm volatile("fma.rn.f32 %0, %1, %2, %1;" : "=r"(d) : "r"(a), "r"(b));                 // FFMA
	if (OP == 18) d = __shfl_xor_sync(0xffffffffu, a, 1) + b;                                               // SHFL (+add)
	if (OP == 19) d = __match_any_sync(0xffffffffu, a & 7) + a;                                             // MATCH.ANY
	if (OP == 20) d = __reduce_or_sync(0xffffffffu, a) + b;                                                 // REDUX.OR
	if (OP == 21) d = __ballot_sync(0xffffffffu, a & 1) + a;                                                // VOTE
	if (OP == 22) d = __funnelshift_l(a, b, a);                                                             // SHF.L.W
	if (OP == 23) d = __clz(a) + b;                                                                         // FLO
	if (OP == 24) d = __popc(a) + b;                                                                        // POPC
	if (OP == 25) asm volatile("bfe.s32 %0, %1, 8, 8;" : "=r"(d) : "r"(a));                                // BFE / SGXT
	if (OP == 26) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=r"(d) : "r"(a));                              // I2F
	if (OP == 27) asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(d) : "r"(a));                             // F2I
	if (OP == 28) d = __vsub2(a, b);                                                                        // VSUB 16x2 (emulated?)
	if (OP == 29) d = __reduce_max_sync(0xffffffffu, a) + b;                                                // REDUX.MAX
	if (OP == 30) asm volatile("shf.r.clamp.b32 %0, %1, %2, %1;" : "=r"(d) : "r"(a), "r"(b));             // SHF.R
	if (OP == 31) asm volatile("mul.lo.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));                      // IMUL
	return d;
}

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, long long* cycles)
{
	uint32_t a[CHAINS];
#pragma unroll
	for (int i = 0; i < CHAINS; i++)
		a[i] = seed * (threadIdx.x + 1) + i;
	const uint32_t b = seed | 3;
	long long t0 = clock64();
#pragma unroll 16
	for (int it = 0; it < ITERS; it++)
	{
#pragma unroll
		for (int i = 0; i < CHAINS; i++)
			a[i] = op<OP>(a[i], b);
	}
	long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < CHAINS; i++)
		s ^= a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0)
		cycles[blockIdx.x] = t1 - t0;
}

// ALU + FMA mix: alternate PRMT (alu) and IMAD (fma) on independent chains
__global__ void __launch_bounds__(256) kmix(uint32_t* out, uint32_t seed, long long* cycles)
{
	uint32_t a[CHAINS];
#pragma unroll
	for (int i = 0; i < CHAINS; i++)
		a[i] = seed * (threadIdx.x + 1) + i;
	const uint32_t b = seed | 3;
	long long t0 = clock64();
#pragma unroll 16
	for (int it = 0; it < ITERS; it++)
	{
#pragma unroll
		for (int i = 0; i < CHAINS; i += 2)
		{
			a[i] = op<4>(a[i], b);
			a[i + 1] = op<0>(a[i + 1], b);
		}
	}
	long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < CHAINS; i++)
		s ^= a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0)
		cycles[blockIdx.x] = t1 - t0;
}

// two ops interleaved on independent chains: do they share a pipe (rate of one) or not (rate of both)?
template <int A, int B>
__global__ void __launch_bounds__(256) kmix2(uint32_t* out, uint32_t seed, long long* cycles)
{
	uint32_t a[CHAINS];
#pragma unroll
	for (int i = 0; i < CHAINS; i++)
		a[i] = seed * (threadIdx.x + 1) + i;
	const uint32_t b = seed | 3;
	long long t0 = clock64();
#pragma unroll 16
	for (int it = 0; it < ITERS; it++)
	{
#pragma unroll
		for (int i = 0; i < CHAINS; i += 2)
		{
			a[i] = op<A>(a[i], b);
			a[i + 1] = op<B>(a[i + 1], b);
		}
	}
	long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < CHAINS; i++)
		s ^= a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0)
		cycles[blockIdx.x] = t1 - t0;
}

template <int A, int B>
void run2(const char* name, uint32_t* d_out, long long* d_cyc)
{
	const int blocks = 148 * 4, threads = 256;
	kmix2<A, B><<<blocks, threads>>>(d_out, 12345, d_cyc);
	cudaDeviceSynchronize();
	kmix2<A, B><<<blocks, threads>>>(d_out, 12345, d_cyc);
	cudaDeviceSynchronize();
	long long h[148 * 4];
	cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < blocks; i++)
		avg += (double)h[i];
	avg /= blocks;
	printf("%-28s %8.1f thread-ops/clk/SM   (%.0f cycles)\n", name, 4.0 * threads * ITERS * CHAINS / avg, avg);
}

template <int OP>
void run(const char* name, uint32_t* d_out, long long* d_cyc)
{
	const int blocks = 148 * 4, threads = 256; // 4 CTAs x 8 warps = 32 warps per SM
	k<OP><<<blocks, threads>>>(d_out, 12345, d_cyc);
	cudaDeviceSynchronize();
	k<OP><<<blocks, threads>>>(d_out, 12345, d_cyc);
	cudaDeviceSynchronize();
	long long h[148 * 4];
	cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0;
	for (int i = 0; i < blocks; i++)
		avg += (double)h[i];
	avg /= blocks;
	// per SM: 4 CTAs x 256 threads x ITERS x CHAINS ops in ~avg cycles (all CTAs resident together)
	printf("%-28s %8.1f thread-ops/clk/SM   (%.0f cycles)\n", name, 4.0 * threads * ITERS * CHAINS / avg, avg);
}

int main()
{
	uint32_t* d_out;
	long long* d_cyc;
	cudaMalloc(&d_out, 148 * 4 * 256 * 4);
	cudaMalloc(&d_cyc, 148 * 4 * 8);
	run<0>("IMAD (mad.lo)", d_out, d_cyc);
	run<1>("IMAD.HI (mul.hi.s32)", d_out, d_cyc);
	run<10>("IMAD.HI + addend (mad.hi)", d_out, d_cyc);
	run<13>("IMAD.WIDE", d_out, d_cyc);
	run<2>("IDP.2A (dp2a)", d_out, d_cyc);
	run<3>("IDP.4A (dp4a)", d_out, d_cyc);
	run<4>("PRMT", d_out, d_cyc);
	run<5>("shr+add (LEA.HI.SX32?)", d_out, d_cyc);
	run<6>("SHF.R.S32", d_out, d_cyc);
	run<7>("LOP3", d_out, d_cyc);
	run<8>("VIADD.16x2", d_out, d_cyc);
	run<9>("add.s32", d_out, d_cyc);
	run<11>("cvt.s32.s16", d_out, d_cyc);
	run<12>("VIMNMX.S16x2", d_out, d_cyc);
	run<14>("viaddmax_s16x2 (DPX)", d_out, d_cyc);
	run<15>("vimax3_s16x2 (DPX)", d_out, d_cyc);
	run<28>("vsub2", d_out, d_cyc);
	run<16>("HFMA2", d_out, d_cyc);
	run<17>("FFMA", d_out, d_cyc);
	run<18>("SHFL.BFLY + add", d_out, d_cyc);
	run<19>("MATCH.ANY + add", d_out, d_cyc);
	run<20>("REDUX.OR + add", d_out, d_cyc);
	run<29>("REDUX.MAX + add", d_out, d_cyc);
	run<21>("VOTE.ballot + add", d_out, d_cyc);
	run<22>("SHF.L.W funnel", d_out, d_cyc);
	run<30>("SHF.R clamp", d_out, d_cyc);
	run<23>("FLO(clz) + add", d_out, d_cyc);
	run<24>("POPC + add", d_out, d_cyc);
	run<25>("BFE.s32", d_out, d_cyc);
	run<26>("I2F", d_out, d_cyc);
	run<27>("F2I", d_out, d_cyc);
	run<31>("mul.lo", d_out, d_cyc);
	{
		const int blocks = 148 * 4, threads = 256;
		kmix<<<blocks, threads>>>(d_out, 12345, d_cyc);
		cudaDeviceSynchronize();
		kmix<<<blocks, threads>>>(d_out, 12345, d_cyc);
		cudaDeviceSynchronize();
		long long h[148 * 4];
		cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
		double avg = 0;
		for (int i = 0; i < blocks; i++)
			avg += (double)h[i];
		avg /= blocks;
		printf("%-28s %8.1f thread-ops/clk/SM   (%.0f cycles)\n", "PRMT+IMAD interleaved", 4.0 * threads * ITERS * CHAINS / avg, avg);
	}
	run2<2, 4>("IDP.2A + PRMT", d_out, d_cyc);
	run2<2, 0>("IDP.2A + IMAD", d_out, d_cyc);
	run2<4, 7>("PRMT + LOP3", d_out, d_cyc);
	run2<5, 0>("shr+add + IMAD", d_out, d_cyc);
	run2<6, 0>("SHF + IMAD", d_out, d_cyc);
	run2<9, 4>("add.s32 + PRMT", d_out, d_cyc);
	return 0;
}
