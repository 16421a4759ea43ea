// unlift_strip.cuh -- the fast inverse-lifting kernel, mirror image of lift_strip.cuh (same strip marching, same
// arithmetic style: dp2a tap sums, hi-domain results, LEA.HI-fused truncation -- see the header of that file).
// Same results as k_unlift_level (lift.cuh), which stays the general kernel.
//
// A CTA owns 120 coefficient columns (+4 halo each side = 128 processed, 136 staged) and marches down the
// level, 8 coefficient rows (16 output rows) per step:
//   * loads: the 8 rows of LL and of the C, B, D subbands of the NEXT step are fetched by the TMA engine, one
//     cp.async.bulk per row, completion on an mbarrier (double buffered). The subbands sit at arbitrary int16
//     offsets of the coefficient stream, so the copy starts at the 16-byte boundary below the first wanted
//     element and the V pass reads with the per-CTA shift (whole words by address, an odd element by a funnel
//     shift). Requires half width % 8 == 0; other shapes use unlift_strip_v1.cuh;
//   * V pass first (lifting.c:118-129): one thread per pair of adjacent columns of one side (left = LL/C,
//     right = B/D); sliding windows in registers for the whole strip height. Inverse quantisation
//     (lifting.c:30-40) is a multiply in the hi domain, whose wrap is the int16 wrap. Even and odd rows go to
//     shared memory as [left 128 | right 128]; the even row of a step is one (DD137) or zero rows ahead of
//     the odd row, rows are independent in the H pass so each carries its own output row number;
//   * H pass (lifting.c:131-133): one thread per (row, 8 coefficients): four 128-bit shared loads, two
//     128-bit global stores of interleaved samples.
// Boundary rules (CLAMP): rows are clamped by the loader's row index; columns outside the level hold garbage
// after the V pass (columns are independent there) and the H pass overrides them in registers:
// H(-1)=H(-2)=H(0), H(t)=H(t-1), E(-1)=E(0), E(t)=E(t+1)=E(t-1). A last odd row / column dropped by the
// plus-one rule is simply not written.
#pragma once

#include "lift_strip.cuh"
#include "unlift_strip_v1.cuh"

constexpr int UT_TW = 120;           // coefficient columns a CTA produces
constexpr int UT_STEP = 8;           // coefficient rows per step
constexpr int UT_THREADS = 128;
constexpr int UT_SP = 152;           // staged row pitch (elements): 136 columns + up to 7 of shift, 16-byte multiple
constexpr int UT_LLW = 136;          // staged LL columns  [c0-8, c0+128)
constexpr int UT_HPW = 144;          // staged subband elements (window starts up to 7 elements early)
constexpr int UT_VP = 264;           // V-pass output row pitch: [left 128 | right 128] + pad, 132 words
constexpr int UT_SBUF = 4 * UT_STEP * UT_SP; // one stage: LL, C, B, D rows

template <int WL>
struct UnstripGeom
{
	static constexpr int LAT = StripGeom<WL>::LAT;              // the odd row a step finishes is this far behind its input row
	static constexpr int LEV = (WL == AKOD_DD137) ? 1 : 0;      // ... and the even row this far
};

// ------------------------------------------------------------------------------------------------
// V pass: per thread two adjacent columns (a = low half, b = high half of every packed word).
// row() consumes lowpass word wl = L(j) and highpass word wh = H(j) (already inverse-quantised) and returns
// the packed even row E(j - LEV) and odd row O(j - LAT).

template <int WL>
struct UnstripV;

template <>
struct UnstripV<AKOD_HAAR>
{
	__device__ __forceinline__ void init() {}
	template <bool EDGE>
	__device__ __forceinline__ void row(uint32_t wl, uint32_t wh, int, int, uint32_t& even, uint32_t& odd)
	{
		even = wl;
		odd = pair_hi((uint32_t)(hd_lo(wl) + hd_lo(wh)), (uint32_t)(hd_hi(wl) + hd_hi(wh))); // O = E + H
	}
};

template <>
struct UnstripV<AKOD_CDF53>
{
	uint32_t wh1;  // H(j-1)
	int ea1, eb1;  // hi-domain E(j-1)
	__device__ __forceinline__ void init()
	{
		wh1 = 0;
		ea1 = eb1 = 0;
	}
	template <bool EDGE>
	__device__ __forceinline__ void row(uint32_t wl, uint32_t wh, int j, int hh, uint32_t& even, uint32_t& odd)
	{
		constexpr uint32_t KM = dpw(-1, -1, 0, 0), KP = dpw(1, 1, 0, 0);
		// E(j) = L(j) - (H(j-1) + H(j)) / 4
		int ea = hi_step<2>(dp2_lo(pair_lo(wh1, wh), KM, 0), hd_lo(wl));
		int eb = hi_step<2>(dp2_lo(pair_hi(wh1, wh), KM, 0), hd_hi(wl));
		if (EDGE && j >= hh)
		{
			ea = ea1; // E(t) = E(t-1)
			eb = eb1;
		}
		// O(j-1) = H(j-1) + (E(j-1) + E(j)) / 2
		const int oa = hi_step<1>(dp2_lo(pair_hi((uint32_t)ea1, (uint32_t)ea), KP, 0), hd_lo(wh1));
		const int ob = hi_step<1>(dp2_lo(pair_hi((uint32_t)eb1, (uint32_t)eb), KP, 0), hd_hi(wh1));
		even = pair_hi((uint32_t)ea, (uint32_t)eb);
		odd = pair_hi((uint32_t)oa, (uint32_t)ob);
		ea1 = ea;
		eb1 = eb;
		wh1 = wh;
	}
};

template <>
struct UnstripV<AKOD_DD137>
{
	uint32_t wh1, wh2, wh3;          // H(j-1), H(j-2), H(j-3)
	uint32_t wl1;                    // L(j-1)
	uint32_t tha2, tha3, thb2, thb3; // th?2 = (H(j-2), H(j-1)), th?3 = (H(j-3), H(j-2)) per column
	int ea2, eb2;                    // hi-domain E(j-2)
	uint32_t tea3, tea4, teb3, teb4; // te?3 = (E(j-3), E(j-2)), te?4 = (E(j-4), E(j-3))
	__device__ __forceinline__ void init()
	{
		wh1 = wh2 = wh3 = wl1 = tha2 = tha3 = thb2 = thb3 = tea3 = tea4 = teb3 = teb4 = 0;
		ea2 = eb2 = 0;
	}
	template <bool EDGE>
	__device__ __forceinline__ void row(uint32_t wl, uint32_t wh, int j, int hh, uint32_t& even, uint32_t& odd)
	{
		constexpr uint32_t KH = dpw(1, -9, -9, 1), KL = dpw(-1, 9, 9, -1);
		const uint32_t tha1 = pair_lo(wh1, wh), thb1 = pair_hi(wh1, wh); // (H(j-1), H(j))
		// E(j-1) = L(j-1) - (-H(j-3) + 9 H(j-2) + 9 H(j-1) - H(j)) / 32
		int ea1 = hi_step<5>(dp2_hi(tha1, KH, dp2_lo(tha3, KH, 0)), hd_lo(wl1));
		int eb1 = hi_step<5>(dp2_hi(thb1, KH, dp2_lo(thb3, KH, 0)), hd_hi(wl1));
		if (EDGE)
		{
			if (j - 1 == 0)
			{
				// E(-1) = E(0): everything older than E(0) becomes E(0)
				ea2 = ea1;
				eb2 = eb1;
				tea3 = tea4 = pair_hi((uint32_t)ea1, (uint32_t)ea1);
				teb3 = teb4 = pair_hi((uint32_t)eb1, (uint32_t)eb1);
			}
			if (j - 1 >= hh)
			{
				ea1 = ea2; // E(t) = E(t+1) = E(t-1)
				eb1 = eb2;
			}
		}
		const uint32_t tea2 = pair_hi((uint32_t)ea2, (uint32_t)ea1), teb2 = pair_hi((uint32_t)eb2, (uint32_t)eb1); // (E(j-2), E(j-1))
		// O(j-3) = H(j-3) - (E(j-4) - 9 E(j-3) - 9 E(j-2) + E(j-1)) / 16
		const int oa = hi_step<4>(dp2_hi(tea2, KL, dp2_lo(tea4, KL, 0)), hd_lo(wh3));
		const int ob = hi_step<4>(dp2_hi(teb2, KL, dp2_lo(teb4, KL, 0)), hd_hi(wh3));
		even = pair_hi((uint32_t)ea1, (uint32_t)eb1);
		odd = pair_hi((uint32_t)oa, (uint32_t)ob);
		tea4 = tea3;
		teb4 = teb3;
		tea3 = tea2;
		teb3 = teb2;
		ea2 = ea1;
		eb2 = eb1;
		tha3 = tha2;
		thb3 = thb2;
		tha2 = tha1;
		thb2 = thb1;
		wh3 = wh2;
		wh2 = wh1;
		wh1 = wh;
		wl1 = wl;
	}
};

// ------------------------------------------------------------------------------------------------
// H pass: lw[8] / hw[8] are the packed lowpass / highpass values of columns c = base - 4 + k, k = 0..15, of one
// row; produces the 8 interleaved (even, odd) sample pairs of columns k = 4..11.
// rem = columns left in the level counted from this chunk's first column (a multiple of 8).

template <int WL>
__device__ __forceinline__ void unstrip_hpass(const uint32_t (&lw)[8], uint32_t (&hw)[8], bool left_edge, int rem,
                                              uint32_t (&out)[8])
{
	// lowpass of column k in the hi domain
	auto lbase = [&](int k) { return (k & 1) ? hd_hi(lw[k >> 1]) : hd_lo(lw[k >> 1]); };
	auto hbase = [&](int k) { return (k & 1) ? hd_hi(hw[k >> 1]) : hd_lo(hw[k >> 1]); };

	if (WL == AKOD_HAAR)
	{
#pragma unroll
		for (int k = 4; k < 12; k++)
		{
			const int e = lbase(k);
			out[k - 4] = pair_hi((uint32_t)e, (uint32_t)(e + hbase(k)));
		}
		return;
	}

	// highpass clamps on the inputs: H(-1) = H(-2) = H(0) (column 0 is k = 4), H(t) = H(t-1) (column t is k = rem + 4,
	// hw[m] holds H(2m), H(2m+1)). The chunk computes E up to k = 13 from H up to k = 14, so the rule also reaches the
	// chunk before a right edge that lies one or two columns into the next chunk (rem = 9, 10)
	if (left_edge)
		hw[1] = pair_lo(hw[2], hw[2]);
	if (rem <= 10)
	{
#pragma unroll
		for (int k = 5; k <= 14; k++)
		{
			// (bit selects on static indices keep hw[] in registers)
			const uint32_t take = 0u - (uint32_t)(rem + 4 == k);
			const uint32_t v = (k & 1) ? pair_lo(hw[k >> 1], hw[k >> 1]) : pair_hi(hw[(k >> 1) - 1], hw[k >> 1]);
			hw[k >> 1] = (v & take) | (hw[k >> 1] & ~take);
		}
	}

	int e[16];
	if (WL == AKOD_CDF53)
	{
		// E(k) = L(k) - (H(k-1) + H(k)) / 4                                       k = 4..12
		constexpr uint32_t KM = dpw(-1, -1, 0, 0);
#pragma unroll
		for (int k = 4; k <= 12; k++)
		{
			int n;
			if (k & 1)
				n = dp2_lo(hw[k >> 1], KM, 0);
			else
				n = dp2_lo(__byte_perm(hw[(k >> 1) - 1], hw[k >> 1], 0x5432), KM, 0); // (H(k-1), H(k))
			e[k] = hi_step<2>(n, lbase(k));
		}
		if (rem <= 8) // E(t) = E(t-1)
		{
#pragma unroll
			for (int k = 12; k >= 5; k--)
			{
				const int take = -(int)(rem + 4 == k);
				e[k] = (e[k - 1] & take) | (e[k] & ~take);
			}
		}
		// O(k) = H(k) + (E(k) + E(k+1)) / 2                                       k = 4..11
		constexpr uint32_t KP = dpw(1, 1, 0, 0);
#pragma unroll
		for (int k = 4; k < 12; k++)
		{
			const int o = hi_step<1>(dp2_lo(pair_hi((uint32_t)e[k], (uint32_t)e[k + 1]), KP, 0), hbase(k));
			out[k - 4] = pair_hi((uint32_t)e[k], (uint32_t)o);
		}
	}
	else
	{
		// E(k) = L(k) - (-H(k-2) + 9 H(k-1) + 9 H(k) - H(k+1)) / 32               k = 3..13
		constexpr uint32_t KH = dpw(1, -9, -9, 1);
		uint32_t ho[8]; // ho[m] = (H(2m-1), H(2m))
#pragma unroll
		for (int m = 1; m <= 7; m++)
			ho[m] = __byte_perm(hw[m - 1], hw[m], 0x5432);
#pragma unroll
		for (int k = 3; k <= 13; k++)
		{
			int n;
			if (k & 1) // taps (k-2, k-1) = ho[(k-1)/2], (k, k+1) = ho[(k+1)/2]
				n = dp2_hi(ho[(k + 1) >> 1], KH, dp2_lo(ho[(k - 1) >> 1], KH, 0));
			else       // taps (k-2, k-1) = hw[k/2 - 1], (k, k+1) = hw[k/2]
				n = dp2_hi(hw[k >> 1], KH, dp2_lo(hw[(k >> 1) - 1], KH, 0));
			e[k] = hi_step<5>(n, lbase(k));
		}
		if (left_edge)
			e[3] = e[4]; // E(-1) = E(0)
		if (rem <= 9) // E(t) = E(t+1) = E(t-1); E(13) is column t when the edge lies one column into the next chunk
		{
#pragma unroll
			for (int k = 13; k >= 5; k--)
			{
				const int take1 = -(int)(rem + 4 == k), take2 = (k >= 6) ? -(int)(rem + 5 == k) : 0;
				e[k] = (e[k - 1] & take1) | ((k >= 6 ? e[k - 2] : 0) & take2) | (e[k] & ~(take1 | take2));
			}
		}
		// O(k) = H(k) - (E(k-1) - 9 E(k) - 9 E(k+1) + E(k+2)) / 16                k = 4..11
		constexpr uint32_t KL = dpw(-1, 9, 9, -1);
		uint32_t te[16]; // te[k] = (E(k), E(k+1))
#pragma unroll
		for (int k = 3; k <= 12; k++)
			te[k] = pair_hi((uint32_t)e[k], (uint32_t)e[k + 1]);
#pragma unroll
		for (int k = 4; k < 12; k++)
		{
			const int o = hi_step<4>(dp2_hi(te[k + 1], KL, dp2_lo(te[k - 1], KL, 0)), hbase(k));
			out[k - 4] = pair_hi((uint32_t)e[k], (uint32_t)o);
		}
	}
}

// ------------------------------------------------------------------------------------------------

template <int WL>
__global__ void __launch_bounds__(UT_THREADS, 6) k_unlift_strip(const UnstripParams up)
{
	constexpr int LAT = UnstripGeom<WL>::LAT, LEV = UnstripGeom<WL>::LEV;
	const UnliftParams& p = up.p;

	__shared__ __align__(16) int16_t S[2 * UT_SBUF];            // staged LL, C, B, D rows, two stages
	__shared__ __align__(16) int16_t VB[2 * 2 * UT_STEP * UT_VP]; // vertically reconstructed rows, double buffered
	__shared__ __align__(8) uint64_t bars[2];

	const int tid = threadIdx.x;
	const uint32_t img = blockIdx.z / p.channels, chn = blockIdx.z - img * p.channels;
	const int hw = (int)p.hw, hh = (int)p.hh;
	const int c0 = blockIdx.x * UT_TW;
	const int i_begin = blockIdx.y * (int)up.split;
	const int i_end = min(i_begin + (int)up.split, hh);
	const uint32_t band = p.hw * p.hh;
	const int16_t* __restrict__ in_ll = p.ll + p.ll_is * img + p.ll_ps * chn;
	const int16_t* __restrict__ in_c = p.stream + p.stream_is * img + p.off_c[chn];
	const int q = (int)__ldg(in_c - 1); // lift head: the decoder learns q from the stream (misc.c:262-268)
	int16_t* __restrict__ out = p.out + p.out_is * img + p.out_ps * chn;

	// ---- loader geometry. Staged column index x <-> level column c0 - 8 + x. A subband row starts at element
	// off_c + band*b + j*hw of the stream, anywhere against 16 bytes: the copy starts at the 16-byte boundary below the
	// first wanted element and the V pass reads with the row's shift (0..7 elements). With hw % 8 == 0 (and then
	// band % 8 == 0) the shift is the same for every row of the three subbands (ALIGNED); otherwise it changes from
	// row to row and band to band and is recomputed where it is needed. Copies take whole 16-byte groups up to the end
	// of the strip or of the row: what they bring in beyond the level's last column is garbage that the H pass
	// overrides (see above), and the rows of the LL plane are padded to 8 elements.
	const uint64_t base_el = p.stream_is * img + p.off_c[chn]; // element offset of the channel's first subband
	const bool aligned = (hw & 7) == 0;
	const int hwm = hw & 7;
	const int sh = (int)(base_el & 7);                          // ALIGNED: the shift of every subband row
	const int x0 = (c0 == 0) ? 8 : 0;                        // first staged column that exists (left edge: columns < 0 do not)
	const int col_end = min(UT_LLW, hw - (c0 - 8));          // one past the last staged column the level has
	const uint32_t ll_bytes = (uint32_t)(((col_end - x0) + 7) & ~7) * 2;

	// Warp b issues the 8 row copies of band b (0 = LL, 1..3 = C, B, D) and arrives on the barrier with that band's
	// bytes: with one issuing warp that warp reached the step's __syncthreads late, every step, and the others waited.
	auto issue = [&](int js, int buf) {
		int16_t* dstbuf = S + buf * UT_SBUF;
		const int lane = tid & 31, b = tid >> 5;
		if (lane < UT_STEP)
		{
			const int r = lane;
			const int j = min(max(js + r, 0), hh - 1);
			int16_t* dst = dstbuf + (b * UT_STEP + r) * UT_SP + x0;
			const int16_t* src;
			uint32_t bytes;
			if (b == 0)
			{
				src = in_ll + (uint32_t)(j * (int)p.ll_rs) + (c0 - 8 + x0);
				bytes = ll_bytes;
			}
			else
			{
				const uint64_t row_el = base_el + (uint64_t)(b - 1) * band + (uint32_t)(j * hw);
				const int shr = (int)(row_el & 7);
				src = p.stream + (row_el - shr) + (c0 - 8 + x0);
				bytes = (uint32_t)(((col_end - x0 + shr) + 7) & ~7) * 2;
			}
			const uint32_t total = __reduce_add_sync((1u << UT_STEP) - 1u, bytes);
			if (lane == 0)
				mbar_expect_tx(&bars[buf], total);
			__syncwarp((1u << UT_STEP) - 1u);
			bulk_g2s(dst, src, bytes, &bars[buf]);
		}
	};
	static_assert(UT_THREADS / 32 == 4, "one issuing warp per band");

	if (tid == 0)
	{
		mbar_init(&bars[0], UT_THREADS / 32); // one arrival (with its expect_tx) per warp
		mbar_init(&bars[1], UT_THREADS / 32);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	// ---- V-pass geometry: thread = (side, column pair); staged columns 4 + 2t, 5 + 2t
	const bool right_side = tid >= 64;
	const int vt = tid & 63;
	const int lo_band = right_side ? 2 : 0, hi_band = right_side ? 3 : 1;   // (LL, C) or (B, D)
	const int hp_word = 2 + vt + (sh >> 1);                                 // ALIGNED: word of the pair in a shifted subband row
	const bool odd_shift = (sh & 1) != 0;
	// not ALIGNED: shift of row j of a band = (band_sh + j * hwm) & 7
	const int hi_sh0 = (int)((base_el + (uint64_t)(hi_band - 1) * band) & 7);
	const int lo_sh0 = right_side ? (int)((base_el + (uint64_t)(lo_band - 1) * band) & 7) : 0;
	const int qhi = q << 16;
	UnstripV<WL> vs;
	vs.init();

	const int j_first = i_begin - LAT, j_last = i_end + LAT;
	issue(j_first, 0);

	int buf = 0;
	uint32_t phase = 0;
	for (int js = j_first; js < j_last; js += UT_STEP)
	{
		int16_t* const VBs = VB + buf * (2 * UT_STEP * UT_VP);
		if (js + UT_STEP < j_last)
			issue(js + UT_STEP, buf ^ 1);
		mbar_wait(&bars[buf], phase);

		// ---------------- V pass
		{
			const uint32_t* sb = reinterpret_cast<const uint32_t*>(S + buf * UT_SBUF);
			const uint32_t* plo = sb + (lo_band * UT_STEP) * (UT_SP / 2) + (right_side ? hp_word : 2 + vt);
			const uint32_t* phi = sb + (hi_band * UT_STEP) * (UT_SP / 2) + hp_word;
			uint32_t* vb = reinterpret_cast<uint32_t*>(VBs) + tid;
			const bool interior = (js - LAT > 0) && (js + UT_STEP <= hh);

			auto vstep = [&](auto edge_tag, auto quant_tag, auto aligned_tag) {
				constexpr bool EDGE = decltype(edge_tag)::value;
				constexpr bool QUANT = decltype(quant_tag)::value;
				constexpr bool ALIGNED = decltype(aligned_tag)::value;
#pragma unroll
				for (int k = 0; k < UT_STEP; k++)
				{
					uint32_t wh, wl;
					if (ALIGNED)
					{
						// odd_shift is uniform per CTA
						wh = phi[k * (UT_SP / 2)];
						if (odd_shift)
							wh = __funnelshift_r(wh, phi[k * (UT_SP / 2) + 1], 16);
						wl = plo[k * (UT_SP / 2)];
						if (odd_shift && right_side)
							wl = __funnelshift_r(wl, plo[k * (UT_SP / 2) + 1], 16);
					}
					else
					{
						// the row's own shift: rows of a level whose width is no multiple of 8 start anywhere
						const int j = EDGE ? min(max(js + k, 0), hh - 1) : js + k;
						const int s_hi = (hi_sh0 + j * hwm) & 7;
						const uint32_t* ph = sb + ((hi_band * UT_STEP + k) * (UT_SP / 2) + 2 + vt) + (s_hi >> 1);
						wh = __funnelshift_r(ph[0], ph[1], (s_hi & 1) << 4);
						if (right_side)
						{
							const int s_lo = (lo_sh0 + j * hwm) & 7;
							const uint32_t* pl = sb + ((lo_band * UT_STEP + k) * (UT_SP / 2) + 2 + vt) + (s_lo >> 1);
							wl = __funnelshift_r(pl[0], pl[1], (s_lo & 1) << 4);
						}
						else
							wl = plo[k * (UT_SP / 2)];
					}
					if (QUANT)
					{
						// lifting.c:30-40: all three highpass subbands; LL is never quantised.
						// (x << 16) * q wraps exactly like the int16 store of x * q.
						wh = pair_hi(wh * (uint32_t)qhi, (wh & 0xffff0000u) * (uint32_t)q);
						if (right_side)
							wl = pair_hi(wl * (uint32_t)qhi, (wl & 0xffff0000u) * (uint32_t)q);
					}
					uint32_t even, odd;
					vs.template row<EDGE>(wl, wh, js + k, hh, even, odd);
					vb[(2 * k) * (UT_VP / 2)] = even;
					vb[(2 * k + 1) * (UT_VP / 2)] = odd;
				}
			};
			auto vstep_q = [&](auto edge_tag, auto quant_tag) {
				if (aligned)
					vstep(edge_tag, quant_tag, std::true_type{});
				else
					vstep(edge_tag, quant_tag, std::false_type{});
			};
			auto vstep_e = [&](auto edge_tag) {
				if (q > 1)
					vstep_q(edge_tag, std::true_type{});
				else
					vstep_q(edge_tag, std::false_type{});
			};
			if (interior)
				vstep_e(std::false_type{});
			else
				vstep_e(std::true_type{});
		}
		__syncthreads();

		// ---------------- H pass: item = (row, chunk of 8 coefficients); 16 rows x 15 chunks
#pragma unroll
		for (int round = 0; round < 2; round++)
		{
			// A half warp takes one row (its 16th lane idles): every quarter warp then reads 128 contiguous bytes
			// of one VB row, so the 128-bit shared loads are conflict-free whatever the row pitch, and the global
			// stores of a half warp are contiguous.
			const int item = tid + UT_THREADS * round;
			const int r = item >> 4, chunk = item & 15;
			const int a = chunk * 8;
			const int cr = js + (r >> 1) - ((r & 1) ? LAT : LEV); // coefficient row this VB row belongs to
			const uint32_t oy = (uint32_t)(2 * cr + (r & 1));
			if (chunk < UT_TW / 8 && c0 + a < hw && (uint32_t)(cr - i_begin) < (uint32_t)(i_end - i_begin) && oy < p.th)
			{
				// VB columns [a, a+16) hold coefficients c = c0 + a - 4 + k
				// explicit 128-bit loads: left to itself the compiler splits the first one into two 64-bit loads
				// (only part of it feeds the left-edge override), which doubles its shared-memory wavefronts
				const uint4 l0 = lds128(&VBs[r * UT_VP + a]);
				const uint4 l1 = lds128(&VBs[r * UT_VP + a + 8]);
				const uint4 g0 = lds128(&VBs[r * UT_VP + 128 + a]);
				const uint4 g1 = lds128(&VBs[r * UT_VP + 128 + a + 8]);
				const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
				uint32_t gw[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
				uint32_t w[8];
				unstrip_hpass<WL>(lw, gw, c0 + a == 0, hw - (c0 + a), w);
				uint4* dst = reinterpret_cast<uint4*>(out + (uint64_t)oy * p.out_rs + 2 * (c0 + a));
				dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
				// the last chunk of a row may end inside its first eight samples: the other eight would land beyond the
				// row's padding, on the next row
				if (2 * (c0 + a) + 8 < (int)p.tw)
					dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
			}
		}
		// No barrier here: the next step's V pass writes the other VB buffer. S[buf] is reloaded by the issue() of
		// the next step, whose V pass (reading S[buf^1]) every thread enters only after leaving this step's V pass;
		// the loads into S[buf] are issued by warp 0 after IT has left this step's V pass, and all other warps
		// left it before the barrier above.
		buf ^= 1;
		if (buf == 0)
			phase ^= 1u;
	}
}

// Any level from 32 x 8 coefficients whose planes have 16-byte aligned rows. An odd output width (tw = 2 hw - 1) has its
// dropped last sample written into the row's padding, which must be there (out_rs >= 2 hw). Subband copies may read up
// to 15 elements beyond a row's end: the caller's stream buffer carries that slack after its last element.
static inline bool unlift_strip_eligible(const UnliftParams& p)
{
	return p.wrap == AKOD_WRAP_CLAMP && (p.tw == 2 * p.hw || p.tw + 1 == 2 * p.hw) && p.out_rs >= 2 * p.hw && p.hw >= 32 && p.hh >= 8 &&
	       (p.out_rs % 8) == 0 && (p.out_ps % 8) == 0 && (p.out_is % 8) == 0 && ((uintptr_t)p.out % 16) == 0 &&
	       (p.ll_rs % 8) == 0 && (p.ll_ps % 8) == 0 && (p.ll_is % 8) == 0 && ((uintptr_t)p.ll % 16) == 0 &&
	       (p.stream_is % 8) == 0 && ((uintptr_t)p.stream % 16) == 0 && p.off_c[0] >= 16 &&
	       (uint64_t)p.tw * p.th < ((uint64_t)1 << 31);
}
