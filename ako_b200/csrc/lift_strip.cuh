// lift_strip.cuh -- the fast forward-lifting kernel ("strip marching"), used for the large CLAMP-wrapped levels
// (any width from 64 samples, any height from 8 coefficient rows; rows 16-byte aligned) that carry nearly all the
// bytes. Same results as k_lift_level (lift.cuh), which stays the general kernel for every other case (other wrap
// modes, tiny levels).
//
// Structure. A CTA owns a strip of 128 coefficient columns (256 samples + 8 halo each side) and MARCHES down
// it, 16 input rows (8 coefficient rows) per step:
//   * loads: the 16 staged rows of the NEXT step are fetched by the TMA engine, one cp.async.bulk per row
//     (544 contiguous bytes) into the other half of a double buffer, completion on an mbarrier. No thread
//     spends instructions or registers on the copy; only the strips that touch the left/right image edge
//     add a few CLAMP fills;
//   * H pass: one thread per (row, 16-pair chunk), six 128-bit shared loads, results back to shared memory
//     as [L x128 | H x128] rows;
//   * V pass: one thread per pair of adjacent columns for the WHOLE strip height: its sliding window stays in
//     registers from step to step (kept as packed column pairs), so nothing is recomputed and no vertical
//     halo is ever re-loaded. Results go straight to global memory, quantised and gated on the way
//     (lifting.c:154-168), at their final offsets in the coefficient stream.
//
// Arithmetic. ncu showed the first version of this kernel bound by the ALU pipe (PRMT/SHF/LEA/LOP3 for
// unpacking int16 pairs, sign extension and truncating division) with the FMA pipe idle, so the lifting
// steps are written to use both pipes and fewer instructions:
//   * tap sums are dp2a (IDP.2A) dot products of packed int16 pairs with constant int8 weights: the unpack
//     and the sign extension disappear into the instruction, and it runs on the FMA pipe;
//   * results live in the "hi domain" (value << 16): x*2^(16-k) + (sample << 16) is ONE IMAD, its upper half
//     is exactly wrap16(sample + floor(x / 2^k)), and PRMT 0x7632 packs two such results without a shift;
//   * truncation toward zero is floor((n + bias) / 2^k) with bias = 2^k - 1 for negative n. |n| < 2^26 here, so
//     the top k bits of n are exactly that bias: n + (n >>> (32-k)) is ONE LEA.HI.
// Every result is identical to the reference's int-promoted expression narrowed to int16 (see hi_step below).
//
// Boundary rules (CLAMP): the loader clamps row/column indices, which gives E(-1)=E(0), E(t)=E(t+1)=E(t-1)
// and the duplicated last row of odd heights; the highpass overrides H(-1)=H(-2)=H(0) and H(t)=H(t-1) are
// applied where those values are produced (wavelet-dd137.c:76-77, :110-111, :122, :151-152, :163).
#pragma once

#include <type_traits>

#include "lift.cuh"

constexpr int FS_TW = 128;             // coefficient columns per strip
constexpr int FS_STEP = 8;             // coefficient rows per step
constexpr int FS_ROWS = 2 * FS_STEP;   // input rows per step
constexpr int FS_THREADS = 128;
constexpr int FS_XW = 2 * FS_TW + 16;  // staged samples per row (8 halo samples each side)
constexpr int FS_XP = 280;             // X row pitch in elements: 140 words, 140 mod 32 = 12 -> conflict-free LDS.128 by row
constexpr int FS_HP = 264;             // [L x128 | H x128] row pitch: 132 words, 132 mod 32 = 4 -> conflict-free STS.128 by row
constexpr int FS_XBUF = FS_ROWS * FS_XP;
constexpr int FS_STAGES = 2;           // staged input buffers: loads run one step ahead of the arithmetic

template <int WL>
struct StripGeom
{
	static constexpr int LAT = (WL == AKOD_DD137) ? 3 : (WL == AKOD_CDF53) ? 1 : 0; // rows of latency of the V pass
};

struct StripParams
{
	LiftParams p;
	uint32_t split; // coefficient rows per CTA (blockIdx.y)
};

// ------------------------------------------------------------------------------------------------
// arithmetic helpers

__device__ __forceinline__ int sx16(int v) // narrow to int16 by wrap, as every store into an int16_t does
{
	return (int)(short)v;
}

__device__ __forceinline__ int lo16(uint32_t w) // even sample of a staged (even, odd) pair
{
	return sx16((int)w);
}

__device__ __forceinline__ int hi16(uint32_t w) // odd sample
{
	return (int)w >> 16;
}

__device__ __forceinline__ uint32_t pack2(int lo, int hi)
{
	return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410);
}

// four int8 weights for dp2a: .lo uses (b0, b1), .hi uses (b2, b3)
__host__ __device__ constexpr uint32_t dpw(int b0, int b1, int b2, int b3)
{
	return (uint32_t)(b0 & 0xff) | ((uint32_t)(b1 & 0xff) << 8) | ((uint32_t)(b2 & 0xff) << 16) | ((uint32_t)(b3 & 0xff) << 24);
}

// c + a.lo16 * w.b0 + a.hi16 * w.b1 (all signed)
__device__ __forceinline__ int dp2_lo(uint32_t a, uint32_t w, int c)
{
	int d;
	asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(w), "r"(c));
	return d;
}

// c + a.lo16 * w.b2 + a.hi16 * w.b3
__device__ __forceinline__ int dp2_hi(uint32_t a, uint32_t w, int c)
{
	int d;
	asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(w), "r"(c));
	return d;
}

// (a.lo16, b.lo16) and (a.hi16, b.hi16) as packed pairs
__device__ __forceinline__ uint32_t pair_lo(uint32_t a, uint32_t b)
{
	return __byte_perm(a, b, 0x5410);
}

__device__ __forceinline__ uint32_t pair_hi(uint32_t a, uint32_t b)
{
	return __byte_perm(a, b, 0x7632);
}

// hi-domain views of the two int16 halves of a word: value << 16, low half zero
__device__ __forceinline__ int hd_lo(uint32_t w)
{
	return (int)(w << 16);
}

__device__ __forceinline__ int hd_hi(uint32_t w)
{
	return (int)(w & 0xffff0000u);
}

// One lifting step in the hi domain: returns X with X >> 16 == wrap16(s + n / 2^K) where '/' truncates toward
// zero as in C, s = base >> 16 (base has a zero low half) and |n| < 2^(31-K).
//   n2 = n + (n >>> (32-K)) adds 2^K - 1 exactly when n is negative; n2 * 2^(16-K) = (floor(n2 / 2^K) << 16) + r,
//   0 <= r < 2^16, so adding base cannot carry out of the low half and the upper half wraps like an int16 store.
template <int K>
__device__ __forceinline__ int hi_step(int n, int base)
{
	const int n2 = n + (int)((uint32_t)n >> (32 - K));
	return n2 * (1 << (16 - K)) + base;
}

// lifting.c:163 on a hi-domain value: (v < -g || v > g) ? v / q : 0. The truncating division is an exact
// 32-bit multiply: with c = ceil(log2 q), s = 15 + c and m = floor(2^s / q) + 1 <= 2^16, the error
// e = m*q - 2^s is in (0, q], so for 0 <= v < 2^15: v*e < 2^s and floor(v*m / 2^s) = floor(v / q); for
// -2^15 <= v < 0: 0 < |v|*e / (q*2^s) <= 1/q and floor(v*m / 2^s) = ceil(v / q) - 1. Adding the sign bit of v
// gives C's truncating v / q in both cases (akod_lift computes m and s on the host).
struct StripQuant
{
	int q, g;
	uint32_t mul;
	int shift;
};

// q == 1 goes through the same expression (m = 2^15 + 1, s = 15: floor(v * m / 2^15) = v for v >= 0, v - 1 for v < 0):
// a channel's q is uniform per CTA, but a branch on it in the store loop costs two instructions per value.
template <bool GATE>
__device__ __forceinline__ int strip_quant(int x_hi, const StripQuant& sq)
{
	const int v = x_hi >> 16;
	int d = (v * (int)sq.mul) >> sq.shift; // |v| * m <= 2^31, exact in 32 bits
	d += (int)((uint32_t)v >> 31);
	if (GATE)
		d = ((uint32_t)(v + sq.g) > (uint32_t)(2 * sq.g)) ? d : 0;
	return d;
}

// ------------------------------------------------------------------------------------------------
// TMA row loads

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}

// one LDS.128, never split by the compiler
__device__ __forceinline__ uint4 lds128(const void* p)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
	return v;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile("{\n\t"
	             ".reg .pred p;\n\t"
	             "WAIT_%=:\n\t"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	             "@p bra DONE_%=;\n\t"
	             "bra WAIT_%=;\n\t"
	             "DONE_%=:\n\t"
	             "}" ::"r"(smem_u32(bar)),
	             "r"(parity)
	             : "memory");
}

// global -> shared bulk copy (TMA engine, SASS UBLKCP); bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

// ------------------------------------------------------------------------------------------------
// H pass: w[0..24) are the staged (even, odd) pairs c = base - 4 + k of one row; produces the 16 lowpass and
// 16 highpass values of pairs k = 4..19 as packed words lw[8], hw[8].

template <int WL>
__device__ __forceinline__ void strip_hpass(const uint32_t (&w)[24], bool left_edge, int rem, uint32_t (&lw)[8],
                                            uint32_t (&hw)[8])
{
	if (WL == AKOD_HAAR)
	{
#pragma unroll
		for (int i = 0; i < 8; i++)
		{
			const int k = 4 + 2 * i;
			lw[i] = pair_lo(w[k], w[k + 1]);                                                // L = e
			hw[i] = pair_hi((uint32_t)(hd_hi(w[k]) - hd_lo(w[k])), (uint32_t)(hd_hi(w[k + 1]) - hd_lo(w[k + 1]))); // H = o - e
		}
	}
	else if (WL == AKOD_CDF53)
	{
		// H(k) = o(k) - (e(k) + e(k+1)) / 2 = o(k) + (-(e(k) + e(k+1))) / 2        k = 3..19
		int H[24];
#pragma unroll
		for (int k = 3; k <= 19; k++)
		{
			const int n = dp2_lo(pair_lo(w[k], w[k + 1]), dpw(-1, -1, 0, 0), 0);
			H[k] = hi_step<1>(n, hd_hi(w[k]));
		}
		if (left_edge)
			H[3] = H[4]; // H(-1) = H(0)
		// L(k) = e(k) + (H(k-1) + H(k)) / 4                                       k = 4..19
		uint32_t TH[24];
#pragma unroll
		for (int k = 3; k <= 18; k++)
			TH[k] = pair_hi((uint32_t)H[k], (uint32_t)H[k + 1]);
		int L[24];
#pragma unroll
		for (int k = 4; k <= 19; k++)
			L[k] = hi_step<2>(dp2_lo(TH[k - 1], dpw(1, 1, 0, 0), 0), hd_lo(w[k]));
#pragma unroll
		for (int i = 0; i < 8; i++)
		{
			lw[i] = pair_hi((uint32_t)L[4 + 2 * i], (uint32_t)L[5 + 2 * i]);
			hw[i] = TH[4 + 2 * i];
		}
	}
	else
	{
		// H(k) = o(k) + (e(k-1) - 9 e(k) - 9 e(k+1) + e(k+2)) / 16                 k = 2..20
		uint32_t T[24];
#pragma unroll
		for (int k = 1; k <= 21; k++)
			T[k] = pair_lo(w[k], w[k + 1]); // (e(k), e(k+1))
		int H[24];
		constexpr uint32_t KH = dpw(1, -9, -9, 1);
#pragma unroll
		for (int k = 2; k <= 20; k++)
		{
			const int n = dp2_hi(T[k + 1], KH, dp2_lo(T[k - 1], KH, 0));
			H[k] = hi_step<4>(n, hd_hi(w[k]));
		}
		if (left_edge)
			H[2] = H[3] = H[4]; // H(-1) = H(-2) = H(0)
		// H(t) = H(t-1); t is 'rem' columns into the chunk of the right edge (any width: 1 .. 16)
		if (rem <= 16)
		{
			// (bit selects on static indices: the compiler turns a chain of conditional moves into an indexed store,
			// which sends H[] to local memory)
#pragma unroll
			for (int k = 20; k >= 5; k--)
			{
				const int take = -(int)(rem + 4 == k);
				H[k] = (H[k - 1] & take) | (H[k] & ~take);
			}
		}
		// L(k) = e(k) + (-H(k-2) + 9 H(k-1) + 9 H(k) - H(k+1)) / 32                k = 4..19
		uint32_t TH[24];
#pragma unroll
		for (int k = 2; k <= 19; k++)
			TH[k] = pair_hi((uint32_t)H[k], (uint32_t)H[k + 1]); // (H(k), H(k+1))
		constexpr uint32_t KL = dpw(-1, 9, 9, -1);
#pragma unroll
		for (int i = 0; i < 8; i++)
		{
			const int k = 4 + 2 * i;
			const int n0 = dp2_hi(TH[k], KL, dp2_lo(TH[k - 2], KL, 0));
			const int n1 = dp2_hi(TH[k + 1], KL, dp2_lo(TH[k - 1], KL, 0));
			lw[i] = pair_hi((uint32_t)hi_step<5>(n0, hd_lo(w[k])), (uint32_t)hi_step<5>(n1, hd_lo(w[k + 1])));
			hw[i] = TH[k];
		}
	}
}

// ------------------------------------------------------------------------------------------------
// V pass: per thread two adjacent columns (a = low half, b = high half of every packed word), state in registers

template <int WL>
struct StripV;

template <>
struct StripV<AKOD_HAAR>
{
	__device__ __forceinline__ void init() {}
	template <bool EDGE>
	__device__ __forceinline__ void row(uint32_t we, uint32_t wo, int, int, int& la, int& lb, int& ha, int& hb)
	{
		la = hd_lo(we);
		lb = hd_hi(we);
		ha = hd_lo(wo) - la;
		hb = hd_hi(wo) - lb;
	}
};

template <>
struct StripV<AKOD_CDF53>
{
	uint32_t we1, wo1; // rows j-1
	uint32_t wh1;      // packed H(j-2)
	__device__ __forceinline__ void init() { we1 = wo1 = wh1 = 0; }
	// consumes input rows (even, odd) of coefficient row j, produces L(j-1), H(j-1)
	template <bool EDGE>
	__device__ __forceinline__ void row(uint32_t we, uint32_t wo, int j, int, int& la, int& lb, int& ha, int& hb)
	{
		constexpr uint32_t KM = dpw(-1, -1, 0, 0), KS = dpw(1, 0, 0, 1);
		const int na = dp2_lo(pair_lo(we1, we), KM, 0), nb = dp2_lo(pair_hi(we1, we), KM, 0);
		ha = hi_step<1>(na, hd_lo(wo1));
		hb = hi_step<1>(nb, hd_hi(wo1));
		const uint32_t wh = pair_hi((uint32_t)ha, (uint32_t)hb);
		if (EDGE && j - 1 == 0)
			wh1 = wh; // H(-1) = H(0)
		const int ua = dp2_lo(wh, KS, dp2_lo(wh1, KS, 0)), ub = dp2_hi(wh, KS, dp2_hi(wh1, KS, 0));
		la = hi_step<2>(ua, hd_lo(we1));
		lb = hi_step<2>(ub, hd_hi(we1));
		wh1 = wh;
		we1 = we;
		wo1 = wo;
	}
};

template <>
struct StripV<AKOD_DD137>
{
	uint32_t we1, we2, we3; // even rows j-1, j-2, j-3
	uint32_t wo1, wo2;      // odd rows j-1, j-2
	uint32_t ta1, ta2, tb1, tb2; // t?1 = (E(j-2), E(j-1)), t?2 = (E(j-3), E(j-2)) per column
	int ha3, hb3;                // hi-domain H(j-3)
	uint32_t tha1, tha2, thb1, thb2; // th?1 = (H(j-4), H(j-3)), th?2 = (H(j-5), H(j-4))
	__device__ __forceinline__ void init()
	{
		we1 = we2 = we3 = wo1 = wo2 = ta1 = ta2 = tb1 = tb2 = tha1 = tha2 = thb1 = thb2 = 0;
		ha3 = hb3 = 0;
	}
	// consumes the rows of coefficient row j, produces L(j-3), H(j-3); th = number of coefficient rows
	template <bool EDGE>
	__device__ __forceinline__ void row(uint32_t we, uint32_t wo, int j, int th, int& la, int& lb, int& ha, int& hb)
	{
		constexpr uint32_t KH = dpw(1, -9, -9, 1), KL = dpw(-1, 9, 9, -1);
		const uint32_t ta0 = pair_lo(we1, we), tb0 = pair_hi(we1, we); // (E(j-1), E(j))
		// H(j-2) = o(j-2) + (E(j-3) - 9 E(j-2) - 9 E(j-1) + E(j)) / 16
		int hna = hi_step<4>(dp2_hi(ta0, KH, dp2_lo(ta2, KH, 0)), hd_lo(wo2));
		int hnb = hi_step<4>(dp2_hi(tb0, KH, dp2_lo(tb2, KH, 0)), hd_hi(wo2));
		if (EDGE)
		{
			if (j - 2 == 0)
			{
				// H(-1) = H(-2) = H(0)
				ha3 = hna;
				hb3 = hnb;
				tha1 = tha2 = pair_hi((uint32_t)hna, (uint32_t)hna);
				thb1 = thb2 = pair_hi((uint32_t)hnb, (uint32_t)hnb);
			}
			if (j - 2 >= th)
			{
				// H(t) = H(t-1)
				hna = ha3;
				hnb = hb3;
			}
		}
		const uint32_t tha0 = pair_hi((uint32_t)ha3, (uint32_t)hna), thb0 = pair_hi((uint32_t)hb3, (uint32_t)hnb); // (H(j-3), H(j-2))
		// L(j-3) = e(j-3) + (-H(j-5) + 9 H(j-4) + 9 H(j-3) - H(j-2)) / 32
		la = hi_step<5>(dp2_hi(tha0, KL, dp2_lo(tha2, KL, 0)), hd_lo(we3));
		lb = hi_step<5>(dp2_hi(thb0, KL, dp2_lo(thb2, KL, 0)), hd_hi(we3));
		ha = ha3;
		hb = hb3;
		ha3 = hna;
		hb3 = hnb;
		tha2 = tha1;
		thb2 = thb1;
		tha1 = tha0;
		thb1 = thb0;
		ta2 = ta1;
		tb2 = tb1;
		ta1 = ta0;
		tb1 = tb0;
		we3 = we2;
		we2 = we1;
		we1 = we;
		wo2 = wo1;
		wo1 = wo;
	}
};

// ------------------------------------------------------------------------------------------------
// MODE (chosen by the host per launch, so that the kernel holds one copy of its V pass per mode):
//   FS_PLAIN  q == 1 and gate == 0 on every channel: the quantise step is the identity
//   FS_QUANT  some channel has q > 1, no channel needs the gate (|v| <= g < q already quantises to zero)
//   FS_GATE   some channel has g >= q
constexpr int FS_PLAIN = 0, FS_QUANT = 1, FS_GATE = 2;

template <int WL, int MODE>
__global__ void __launch_bounds__(FS_THREADS, 6) k_lift_strip(const StripParams sp)
{
	constexpr bool PLAIN = MODE == FS_PLAIN, GATE = MODE == FS_GATE;
	constexpr int LAT = StripGeom<WL>::LAT;
	const LiftParams& p = sp.p;

	__shared__ __align__(16) int16_t X[FS_STAGES * FS_XBUF];
	__shared__ __align__(16) int16_t HB[2 * FS_ROWS * FS_HP]; // double buffered: one barrier per step
	__shared__ __align__(8) uint64_t bars[FS_STAGES];

	const int tid = threadIdx.x;
	const uint32_t img = blockIdx.z / p.channels, chn = blockIdx.z - img * p.channels;
	const int tw = (int)p.tw, th = (int)p.th, cw = (int)p.cw;
	const int c0 = blockIdx.x * FS_TW;
	const int i_begin = blockIdx.y * (int)sp.split;
	const int i_end = min(i_begin + (int)sp.split, th);
	const int16_t* __restrict__ in = p.in + p.in_is * img + p.in_ps * chn;

	const uint32_t band = p.tw * p.th; // < 2^31 elements (host-checked)
	int16_t* __restrict__ ll = p.ll + p.ll_is * img + p.ll_ps * chn;
	int16_t* __restrict__ out_c = p.stream + p.stream_is * img + p.off_c[chn];
	StripQuant sq;
	sq.q = p.q[chn];
	sq.g = p.g[chn];
	sq.mul = p.qmul[chn];
	sq.shift = p.qshift[chn];
	if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)
		out_c[-1] = (int16_t)sq.q; // akoLiftHead

	// ---- loader geometry: staged sample x of a row is input sample xs0 + x; [xa, xb) is inside the row. Rows start on
	// 16-byte boundaries (in_rs % 8 == 0) and xs0 is a multiple of 8, so the TMA copy takes [xa, xb8), xb8 = xb rounded
	// down to 8 samples; a width that is no multiple of 8 leaves up to 7 samples of the last strip to the fill threads.
	const int xs0 = 2 * c0 - 8;
	const int xa = (xs0 < 0) ? 8 : 0;                       // first staged sample that exists
	const int xb = min(FS_XW, cw - xs0);                    // one past the last
	const int xb8 = xb & ~7;
	const int in_rs = (int)p.in_rs, last_row = (int)p.ch - 1;
	const uint32_t row_bytes = (uint32_t)(xb8 - xa) * 2;
	const bool edge_strip = (xa != 0) || (xb != FS_XW);
	const int x_last_even = (cw - 1) & ~1;                  // CLAMP source on the right (the "fake last" odd sample too)

	// One warp issues the step's row copies. (Spreading them over the four warps, as the inverse kernel does with
	// its 32 copies per step, gained 1 % on 8192^2 planes and lost 2 % on the 816-column C2 planes, where 2 of 7
	// strips take the edge-fill branch below and every warp would then diverge on it.)
	auto issue = [&](int js, int buf) {
		int16_t* dstbuf = X + buf * FS_XBUF;
		if (tid < 32)
		{
			if (tid == 0)
				mbar_expect_tx(&bars[buf], row_bytes * FS_ROWS);
			__syncwarp();
			if (tid < FS_ROWS)
			{
				const int j = min(max(js + (tid >> 1), 0), th - 1);
				const int y = min(2 * j + (tid & 1), last_row);
				bulk_g2s(dstbuf + tid * FS_XP + xa, in + (uint32_t)(y * in_rs) + (xs0 + xa), row_bytes, &bars[buf]);
			}
		}
		else if (edge_strip)
		{
			// CLAMP: every sample outside the row is the first / last even sample, 8 samples per store; the group that
			// straddles the end of a row whose width is no multiple of 8 also carries the row's last samples
			const int nleft = xa >> 3, nright = (FS_XW - xb8) >> 3;
			for (int i = tid - 32; i < FS_ROWS * (nleft + nright); i += FS_THREADS - 32)
			{
				const int r = i / (nleft + nright), v = i - r * (nleft + nright);
				const int j = min(max(js + (r >> 1), 0), th - 1);
				const int y = min(2 * j + (r & 1), last_row);
				const int16_t* row = in + (uint32_t)(y * in_rs);
				const bool left = v < nleft;
				const uint32_t e = (uint16_t)__ldg(row + (left ? 0 : x_last_even));
				const uint32_t w = e * 0x10001u;
				const int x = left ? 8 * v : xb8 + 8 * (v - nleft);
				uint4 q = make_uint4(w, w, w, w);
				if (!left && x < xb)
				{
					uint32_t m[4];
#pragma unroll
					for (int k = 0; k < 4; k++)
					{
						const uint32_t lo = (x + 2 * k < xb) ? (uint32_t)(uint16_t)__ldg(row + xs0 + x + 2 * k) : e;
						const uint32_t hi = (x + 2 * k + 1 < xb) ? (uint32_t)(uint16_t)__ldg(row + xs0 + x + 2 * k + 1) : e;
						m[k] = lo | (hi << 16);
					}
					q = make_uint4(m[0], m[1], m[2], m[3]);
				}
				*reinterpret_cast<uint4*>(dstbuf + r * FS_XP + x) = q;
			}
		}
	};

	if (tid == 0)
	{
#pragma unroll
		for (int i = 0; i < FS_STAGES; i++)
			mbar_init(&bars[i], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	// ---- V-pass thread geometry
	const bool right_half = tid >= FS_TW / 2;                               // H-pass highpass half -> B, D
	const int vcol = c0 + 2 * (right_half ? tid - FS_TW / 2 : tid);         // coefficient column of the pair
	const bool vvalid = vcol < tw;
	// V-high goes to D or C, V-low to B or (left half) the next level's input plane
	int16_t* const out_hi = (right_half ? out_c + 2 * (uint64_t)band : out_c) + vcol;
	int16_t* const out_lo = right_half ? (out_c + band + vcol) : (ll + vcol);
	const uint32_t hi_rs = (uint32_t)tw, lo_rs = right_half ? (uint32_t)tw : p.ll_rs;
	// The C/B/D subbands of a channel start at an odd or even int16 offset of the stream (a 2-byte lift head
	// precedes each channel's block, so the parity alternates from channel to channel): pairs are stored with
	// one 32-bit store when aligned, two 16-bit stores otherwise. Uniform per CTA.
	// ... or the level has an odd number of coefficient columns, so that the parity alternates from row to row; the
	// last thread of such a level owns one column only.
	const bool odd_offset = ((p.off_c[chn] | p.tw) & 1) != 0;
	const bool vsingle = vcol + 1 >= tw;
	StripV<WL> vs;
	vs.init();

	const int j_first = i_begin - LAT;
	const int j_last = i_end + LAT; // exclusive
#pragma unroll
	for (int i = 0; i < FS_STAGES - 1; i++)
		if (j_first + i * FS_STEP < j_last)
			issue(j_first + i * FS_STEP, i);
	__syncthreads(); // edge fills of the first buffers

	int buf = 0, nbuf = FS_STAGES - 1, hb = 0;
	uint32_t phase = 0;
	for (int js = j_first; js < j_last; js += FS_STEP)
	{
		int16_t* const HBs = HB + hb * (FS_ROWS * FS_HP);
		// buffer nbuf was last read by the H pass of the previous step, which every thread has left
		if (js + (FS_STAGES - 1) * FS_STEP < j_last)
			issue(js + (FS_STAGES - 1) * FS_STEP, nbuf);
		mbar_wait(&bars[buf], phase);

		// ---------------- H pass: thread = (row, chunk of 16 coefficient pairs)
		{
			const int r = tid & 15, chunk = tid >> 4;
			const int a = chunk * 16;
			if (c0 + a < tw)
			{
				// words [a, a+24) of the staged row hold pairs c = c0 + a - 4 + k
				uint32_t w[24];
				const uint4* src = reinterpret_cast<const uint4*>(&X[buf * FS_XBUF + r * FS_XP + 2 * a]);
#pragma unroll
				for (int k = 0; k < 6; k++)
				{
					const uint4 t = src[k];
					w[4 * k] = t.x;
					w[4 * k + 1] = t.y;
					w[4 * k + 2] = t.z;
					w[4 * k + 3] = t.w;
				}
				uint32_t lw[8], hw[8];
				strip_hpass<WL>(w, c0 + a == 0, tw - (c0 + a), lw, hw);
				uint4* dl = reinterpret_cast<uint4*>(&HBs[r * FS_HP + a]);
				uint4* dh = reinterpret_cast<uint4*>(&HBs[r * FS_HP + FS_TW + a]);
				dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
				dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
				dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
				dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
			}
		}
		__syncthreads();

		// ---------------- V pass: marching, state in registers
		if (vvalid)
		{
			const uint32_t* col = reinterpret_cast<const uint32_t*>(HBs) + tid;
			const int i0 = js - LAT; // output row of this step's first input row
			// all eight output rows inside [i_begin, i_end) and no boundary rule fires in this step
			const bool interior = (i0 >= i_begin) && (i0 + FS_STEP <= i_end) && (js > LAT) && (js + FS_STEP <= th);
			int16_t* const row_hi = out_hi + (int64_t)i0 * (int64_t)hi_rs;
			int16_t* const row_lo = out_lo + (int64_t)i0 * (int64_t)lo_rs;

			// EDGE: a boundary rule or the CTA's row range cuts this step. ODD: the channel's subbands start at an odd
			// int16 offset of the stream (uniform per CTA; as a tag it costs no branch per row).
			auto vstep = [&](auto edge_tag, auto odd_tag) {
				constexpr bool EDGE = decltype(edge_tag)::value;
				constexpr bool ODD = decltype(odd_tag)::value;
#pragma unroll
				for (int k = 0; k < FS_STEP; k++)
				{
					const uint32_t we = col[(2 * k) * (FS_HP / 2)], wo = col[(2 * k + 1) * (FS_HP / 2)];
					int la, lb, ha, hb;
					vs.template row<EDGE>(we, wo, js + k, th, la, lb, ha, hb);
					if (!EDGE || (uint32_t)(i0 + k - i_begin) < (uint32_t)(i_end - i_begin))
					{
						int16_t* dh = row_hi + (uint32_t)k * hi_rs;
						int16_t* dl = row_lo + (uint32_t)k * lo_rs;
						uint32_t whi, wlo;
						if (PLAIN)
							whi = pair_hi((uint32_t)ha, (uint32_t)hb);
						else
							whi = pack2(strip_quant<GATE>(ha, sq), strip_quant<GATE>(hb, sq));
						if (PLAIN || !right_half)
							wlo = pair_hi((uint32_t)la, (uint32_t)lb);
						else
							wlo = pack2(strip_quant<GATE>(la, sq), strip_quant<GATE>(lb, sq));
						if (ODD)
						{
							dh[0] = (int16_t)whi;
							if (!vsingle)
								dh[1] = (int16_t)(whi >> 16);
						}
						else
							*reinterpret_cast<uint32_t*>(dh) = whi;
						if (ODD && (right_half || vsingle))
						{
							dl[0] = (int16_t)wlo;
							if (!vsingle)
								dl[1] = (int16_t)(wlo >> 16);
						}
						else
							*reinterpret_cast<uint32_t*>(dl) = wlo;
					}
				}
			};
			auto vstep_o = [&](auto edge_tag) {
				if (odd_offset)
					vstep(edge_tag, std::true_type{});
				else
					vstep(edge_tag, std::false_type{});
			};
			if (interior)
				vstep_o(std::false_type{});
			else
				vstep_o(std::true_type{});
		}
		// No barrier here: the next step's H pass writes the other HB buffer; this one is rewritten two steps on,
		// behind the next step's barrier. X[buf] is reloaded by the issue() at the top of the next step -- every
		// thread has left this step's H pass (it is behind the barrier above) by then.
		hb ^= 1;
		nbuf = buf;
		if (++buf == FS_STAGES)
		{
			buf = 0;
			phase ^= 1u;
		}
	}
}

// host-side eligibility of a level for the strip kernel
static inline bool lift_strip_eligible(const LiftParams& p)
{
	return p.wrap == AKOD_WRAP_CLAMP && p.cw >= 64 && p.th >= 8 && (p.in_rs % 8) == 0 &&
	       (p.in_ps % 8) == 0 && (p.in_is % 8) == 0 && ((uintptr_t)p.in % 16) == 0 && (p.ll_rs % 2) == 0 &&
	       (p.ll_ps % 2) == 0 && (p.ll_is % 2) == 0 && ((uintptr_t)p.ll % 4) == 0 && (p.stream_is % 2) == 0 &&
	       ((uintptr_t)p.stream % 4) == 0 && (uint64_t)p.cw * p.ch < ((uint64_t)1 << 31);
}
