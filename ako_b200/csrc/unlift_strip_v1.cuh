// unlift_strip_v1.cuh -- first inverse strip kernel (register-prefetch loader, scalar arithmetic). Kept for levels whose
// half width is a multiple of 4 but not of 8, which the TMA-fed kernel of unlift_strip.cuh cannot take.
// Mirror image of lift_strip.cuh's structure. Same arithmetic as
// k_unlift_level (lift.cuh), which stays the general kernel (other wrap modes, unaligned widths, tiny levels).
//
// A CTA owns 120 coefficient columns (+4 halo each side = 128 staged) and marches down the level, 8
// coefficient rows (16 output rows) per step:
//   * staging: the 8 rows of LL (64-bit loads) and of C, B, D (32-bit loads, or 16-bit when the subband
//     starts at an odd int16 offset of the stream) are prefetched into registers one step ahead;
//   * V pass first (lifting.c:118-129): one thread per pair of adjacent columns of one side (left = LL/C,
//     right = B/D); the sliding windows live in registers for the whole strip height. Inverse quantisation
//     (lifting.c:30-40) is fused here. Even and odd rows go to shared memory as [left 128 | right 128];
//   * H pass (lifting.c:131-133): one thread per (row, 8 coefficients): four 128-bit shared loads, 11 even
//     + 8 odd samples from registers, two 128-bit global stores of interleaved samples.
// Boundary rules (CLAMP): highpass inputs are clamped by the loader (H(-1)=H(-2)=H(0), H(t)=H(t-1)); the
// computed evens are overridden where produced (E(-1)=E(0), E(t)=E(t+1)=E(t-1)); a last odd row / column
// dropped by the plus-one rule is simply not written.
#pragma once

#include "lift_strip.cuh"

constexpr int US_TW = 120;           // coefficient columns a CTA produces
constexpr int US_SW = 128;           // staged columns (4 halo each side)
constexpr int US_STEP = 8;           // coefficient rows per step
constexpr int US_THREADS = 128;
constexpr int US_SP = 136;           // staged row pitch (elements): 68 words, 68 mod 32 = 4
constexpr int US_VP = 264;           // V-pass output row pitch: [left 128 | right 128] + pad, 132 words
constexpr int US_LL_LOADS = (US_STEP * US_SW / 4) / US_THREADS;      // 2  64-bit loads per thread
constexpr int US_HP_LOADS = (3 * US_STEP * US_SW / 2) / US_THREADS + 1; // 12 words per thread + the 65th word of a row

struct UnstripParams
{
	UnliftParams p;
	uint32_t split;
};

template <int WL>
__device__ __forceinline__ int ustrip_even(int lp, int l2, int l1, int h, int p1)
{
	if (WL == AKOD_HAAR)
		return lp;
	if (WL == AKOD_CDF53)
		return sx16(lp - (l1 + h) / 4);
	return sx16(lp - (-l2 - p1 + 9 * (l1 + h)) / 32);
}

template <int WL>
__device__ __forceinline__ int ustrip_odd(int hp, int e, int l1, int p1, int p2)
{
	if (WL == AKOD_HAAR)
		return e + hp;
	if (WL == AKOD_CDF53)
		return hp + (e + p1) / 2;
	return hp - (l1 + p2 - 9 * (e + p1)) / 16;
}

template <int WL>
__global__ void __launch_bounds__(US_THREADS, 6) k_unlift_strip_v1(const UnstripParams up)
{
	constexpr int LAT = StripGeom<WL>::LAT;
	const UnliftParams& p = up.p;

	__shared__ __align__(16) int16_t S[4 * US_STEP * US_SP];  // staged LL, C, B, D rows
	__shared__ __align__(16) int16_t VB[2 * US_STEP * US_VP]; // vertically reconstructed rows

	const int tid = threadIdx.x;
	const uint32_t img = blockIdx.z / p.channels, chn = blockIdx.z - img * p.channels;
	const int hw = (int)p.hw, hh = (int)p.hh;
	const int c0 = blockIdx.x * US_TW;
	const int i_begin = blockIdx.y * (int)up.split;
	const int i_end = min(i_begin + (int)up.split, hh);
	const uint32_t band = p.hw * p.hh;
	const int16_t* __restrict__ in_ll = p.ll + p.ll_is * img + p.ll_ps * chn;
	const int16_t* __restrict__ in_c = p.stream + p.stream_is * img + p.off_c[chn];
	const int q = (int)__ldg(in_c - 1); // lift head: the decoder learns q from the stream (misc.c:262-268)
	const int shift1 = (int)(p.off_c[chn] & 1); // 1: C/B/D start at an odd int16 offset of the stream
	const int fshift = 16 * shift1;
	int16_t* __restrict__ out = p.out + p.out_is * img + p.out_ps * chn;

	// ---- loader
	uint2 pre_ll[US_LL_LOADS];
	uint32_t pre_hp[US_HP_LOADS];
	auto prefetch = [&](int js) {
#pragma unroll
		for (int k = 0; k < US_LL_LOADS; k++)
		{
			const int id = tid + US_THREADS * k;
			const int r = id >> 5, v = id & 31;
			const int j = min(max(js + r, 0), hh - 1);
			const int c = c0 - 4 + 4 * v;
			const int16_t* row = in_ll + (uint32_t)(j * (int)p.ll_rs);
			if (c >= 0 && c + 4 <= hw)
				pre_ll[k] = __ldg(reinterpret_cast<const uint2*>(row + c));
			else
			{
				const uint32_t e = (uint16_t)__ldg(row + (c < 0 ? 0 : hw - 1));
				pre_ll[k] = make_uint2(e * 0x10001u, e * 0x10001u);
			}
		}
		// Always aligned 32-bit loads: when the subband starts at an odd int16 offset the staged row starts one
		// column earlier (c0-5) and the V pass realigns pairs with a funnel shift. 65 words cover the 128 staged
		// columns at either parity: slots 0..11 are words 0..63 of the 24 band rows, slot 12 is word 64 (24 threads).
#pragma unroll
		for (int k = 0; k < US_HP_LOADS; k++)
		{
			int bnd, r, m;
			if (k < US_HP_LOADS - 1)
			{
				const int id = tid + US_THREADS * k;
				bnd = id >> 9, r = (id >> 6) & 7, m = id & 63; // band 0..2 = C, B, D
			}
			else
				bnd = tid >> 3, r = tid & 7, m = 64;
			if (k < US_HP_LOADS - 1 || tid < 3 * US_STEP)
			{
				const int j = min(max(js + r, 0), hh - 1);
				const int c = c0 - 4 - shift1 + 2 * m;
				const int16_t* row = in_c + (uint64_t)bnd * band + (uint32_t)(j * hw);
				if (c >= 0 && c + 2 <= hw)
					pre_hp[k] = __ldg(reinterpret_cast<const uint32_t*>(row + c));
				else
				{
					// CLAMP, element by element (edge strips only)
					const int ca = min(max(c, 0), hw - 1), cb = min(max(c + 1, 0), hw - 1);
					pre_hp[k] = (uint32_t)(uint16_t)__ldg(row + ca) | ((uint32_t)(uint16_t)__ldg(row + cb) << 16);
				}
			}
		}
	};
	auto commit = [&]() {
#pragma unroll
		for (int k = 0; k < US_LL_LOADS; k++)
		{
			const int id = tid + US_THREADS * k;
			*reinterpret_cast<uint2*>(&S[(id >> 5) * US_SP + 4 * (id & 31)]) = pre_ll[k];
		}
#pragma unroll
		for (int k = 0; k < US_HP_LOADS; k++)
		{
			int bnd, r, m;
			if (k < US_HP_LOADS - 1)
			{
				const int id = tid + US_THREADS * k;
				bnd = id >> 9, r = (id >> 6) & 7, m = id & 63;
			}
			else
				bnd = tid >> 3, r = tid & 7, m = 64;
			if (k < US_HP_LOADS - 1 || tid < 3 * US_STEP)
				*reinterpret_cast<uint32_t*>(&S[((bnd + 1) * US_STEP + r) * US_SP + 2 * m]) = pre_hp[k];
		}
	};

	// ---- V-pass state (two adjacent columns of one side per thread)
	const bool right_side = tid >= 64;
	const int vword = tid & 63; // word (column pair) inside the side's 128 staged columns
	const int16_t* s_lo = S + (right_side ? 2 * US_STEP * US_SP : 0);             // LL or B
	const int16_t* s_hi = S + (right_side ? 3 * US_STEP * US_SP : US_STEP * US_SP); // C or D
	int h1[2] = {0, 0}, h2[2] = {0, 0}, h3[2] = {0, 0}, l1[2] = {0, 0};
	int ev2[2] = {0, 0}, ev3[2] = {0, 0}, ev4[2] = {0, 0};

	const int j_first = i_begin - LAT, j_last = i_end + LAT;
	const uint32_t n_out = (uint32_t)(i_end - i_begin);
	prefetch(j_first);

	for (int js = j_first; js < j_last; js += US_STEP)
	{
		commit();
		__syncthreads();
		if (js + US_STEP < j_last)
			prefetch(js + US_STEP);

		// ---------------- V pass
		{
			const uint32_t* plo = reinterpret_cast<const uint32_t*>(s_lo) + vword;
			const uint32_t* phi = reinterpret_cast<const uint32_t*>(s_hi) + vword;
			uint32_t* vb = reinterpret_cast<uint32_t*>(VB) + (right_side ? 64 : 0) + vword;
#pragma unroll
			for (int k = 0; k < US_STEP; k++)
			{
				const int j = js + k;
				// staged highpass rows may be shifted by one column (see the loader): realign with a funnel shift
				const uint32_t wh = __funnelshift_r(phi[k * (US_SP / 2)], phi[k * (US_SP / 2) + 1], fshift);
				const uint32_t wl = right_side ? __funnelshift_r(plo[k * (US_SP / 2)], plo[k * (US_SP / 2) + 1], fshift)
				                               : plo[k * (US_SP / 2)];
				int lv[2] = {lo16(wl), hi16(wl)};
				int hv[2] = {lo16(wh), hi16(wh)};
				if (q > 1)
				{
					// lifting.c:30-40: all three highpasses; LL is never quantised
					hv[0] = sx16(hv[0] * q);
					hv[1] = sx16(hv[1] * q);
					if (right_side)
					{
						lv[0] = sx16(lv[0] * q);
						lv[1] = sx16(lv[1] * q);
					}
				}
				int even[2], odd[2];
#pragma unroll
				for (int s = 0; s < 2; s++)
				{
					if (WL == AKOD_DD137)
					{
						int e = ustrip_even<WL>(l1[s], h3[s], h2[s], h1[s], hv[s]); // even(j-1)
						if (j - 1 == 0)
							ev2[s] = e; // E(-1) = E(0)
						if (j - 1 >= hh)
							e = ev2[s]; // E(t) = E(t+1) = E(t-1)
						odd[s] = ustrip_odd<WL>(h3[s], ev3[s], ev4[s], ev2[s], e); // odd(j-3)
						even[s] = ev3[s];
						ev4[s] = ev3[s];
						ev3[s] = ev2[s];
						ev2[s] = e;
						h3[s] = h2[s];
						h2[s] = h1[s];
						h1[s] = hv[s];
						l1[s] = lv[s];
					}
					else if (WL == AKOD_CDF53)
					{
						int e = ustrip_even<WL>(lv[s], 0, h1[s], hv[s], 0); // even(j)
						if (j >= hh)
							e = ev2[s]; // E(t) = E(t-1)
						odd[s] = ustrip_odd<WL>(h1[s], ev2[s], 0, e, 0); // odd(j-1)
						even[s] = ev2[s];
						ev2[s] = e;
						h1[s] = hv[s];
					}
					else
					{
						even[s] = lv[s];
						odd[s] = ustrip_odd<WL>(hv[s], lv[s], 0, 0, 0);
					}
				}
				vb[(2 * k) * (US_VP / 2)] = pack2(even[0], even[1]);
				vb[(2 * k + 1) * (US_VP / 2)] = pack2(odd[0], odd[1]);
			}
		}
		__syncthreads();

		// ---------------- H pass: item = (row, chunk of 8 coefficients); 16 rows x 15 chunks
		const int r0 = js - LAT; // coefficient row of VB rows 0,1
#pragma unroll
		for (int round = 0; round < 2; round++)
		{
			// consecutive lanes take consecutive chunks of a row: conflict-free 128-bit shared loads and
			// contiguous global stores
			const int item = tid + US_THREADS * round;
			const int r = item / (US_TW / 8), chunk = item - r * (US_TW / 8);
			const int a = chunk * 8;
			const int cr = r0 + (r >> 1); // coefficient row this VB row belongs to
			const uint32_t oy = (uint32_t)(2 * cr + (r & 1));
			if (r < 2 * US_STEP && c0 + a < hw && (uint32_t)(cr - i_begin) < n_out && oy < p.th)
			{
				// VB columns [a, a+16) hold coefficients c = c0 + a - 4 + k
				int L[16], Hc[16];
				{
					const uint4 l0 = *reinterpret_cast<const uint4*>(&VB[r * US_VP + a]);
					const uint4 l1v = *reinterpret_cast<const uint4*>(&VB[r * US_VP + a + 8]);
					const uint4 g0 = *reinterpret_cast<const uint4*>(&VB[r * US_VP + 128 + a]);
					const uint4 g1 = *reinterpret_cast<const uint4*>(&VB[r * US_VP + 128 + a + 8]);
					const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1v.x, l1v.y, l1v.z, l1v.w};
					const uint32_t gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
					for (int k = 0; k < 8; k++)
					{
						L[2 * k] = lo16(lw[k]);
						L[2 * k + 1] = hi16(lw[k]);
						Hc[2 * k] = lo16(gw[k]);
						Hc[2 * k + 1] = hi16(gw[k]);
					}
				}
				int e[16];
#pragma unroll
				for (int k = 3; k <= 13; k++)
				{
					if (WL == AKOD_DD137)
						e[k] = ustrip_even<WL>(L[k], Hc[k - 2], Hc[k - 1], Hc[k], Hc[k + 1]);
					else
						e[k] = ustrip_even<WL>(L[k], 0, Hc[k - 1], Hc[k], 0);
				}
				if (WL != AKOD_HAAR)
				{
					if (c0 + a == 0)
						e[3] = e[4]; // E(-1) = E(0)
					const int rem = hw - (c0 + a); // E(t) = E(t+1) = E(t-1); t is 4 or 8 columns into an edge chunk
					if (rem == 4)
						e[8] = e[9] = e[7];
					if (rem == 8)
						e[12] = e[13] = e[11];
				}
				uint32_t w[8];
#pragma unroll
				for (int k = 4; k < 12; k++)
				{
					int o;
					if (WL == AKOD_DD137)
						o = ustrip_odd<WL>(Hc[k], e[k], e[k - 1], e[k + 1], e[k + 2]);
					else
						o = ustrip_odd<WL>(Hc[k], e[k], 0, e[k + 1], 0);
					w[k - 4] = pack2(e[k], o);
				}
				uint4* dst = reinterpret_cast<uint4*>(out + (uint64_t)oy * p.out_rs + 2 * (c0 + a));
				dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
				if (c0 + a + 4 < hw)
					dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
			}
		}
		__syncthreads();
	}
}

static inline bool unlift_strip_v1_eligible(const UnliftParams& p)
{
	return p.wrap == AKOD_WRAP_CLAMP && (p.tw % 8) == 0 && p.tw == 2 * p.hw && p.hw >= 32 && p.hh >= 8 &&
	       (p.out_rs % 8) == 0 && (p.out_ps % 8) == 0 && (p.out_is % 8) == 0 && ((uintptr_t)p.out % 16) == 0 &&
	       (p.ll_rs % 4) == 0 && (p.ll_ps % 4) == 0 && (p.ll_is % 4) == 0 && ((uintptr_t)p.ll % 8) == 0 &&
	       (p.stream_is % 2) == 0 && ((uintptr_t)p.stream % 4) == 0 && (uint64_t)p.tw * p.th < ((uint64_t)1 << 31);
}
