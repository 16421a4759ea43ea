// lift_strip4.cuh -- level 0 of the forward transform straight from the interleaved RGBA8 image: the colour/format
// pass (format.c:64-135) fused into the strip-marching lifting kernel of lift_strip.cuh.
//
// Unfused, k_format_fwd_rgba8x8 reads 4 B/pixel and writes four int16 planes (8 B/pixel) that the level-0 lifting
// kernel reads right back: 16 B/pixel of HBM traffic and a launch that fusion removes. Here a CTA owns a strip of 64
// coefficient columns of ALL FOUR channels (256 threads, 64 per channel) and marches down it, 16 image rows per step:
//   * loads: the TMA engine fetches the 16 RGBA8 rows of the NEXT step (576 contiguous bytes each) into the other
//     half of a double buffer, completion on an mbarrier;
//   * conversion: every thread converts quads of pixels (one LDS.128) to the four planes' int16 samples -- YCoCg /
//     YCoCg_Q / subtract-G / none, alpha discard, exactly color_forward() of format.cuh -- and stores them into the
//     per-channel staged rows; the CLAMP columns outside the image are filled here as well;
//   * H pass / V pass: per channel, exactly as in k_lift_strip (strip_hpass, StripV: same arithmetic, same boundary
//     rules, same quantise + gate at the store), on 64-column strips: thread = (row, 16-pair chunk) in the H pass,
//     thread = pair of adjacent columns for the whole strip height in the V pass, its window in registers.
// Two CTA-wide barriers per step (conversion -> H, H -> V). Results are identical to k_format_fwd_* followed by
// k_lift_strip.
#pragma once

#include "format.cuh"
#include "lift_strip.cuh"

constexpr int F4_TW = 64;               // coefficient columns per strip
constexpr int F4_CH = 4;
constexpr int F4_GROUP = F4_TW;         // threads per channel
constexpr int F4_THREADS = F4_CH * F4_GROUP;
constexpr int F4_XW = 2 * F4_TW + 16;   // staged samples per row (8 halo samples each side)
constexpr int F4_RAW_ROW = F4_XW * 4;   // bytes of one staged RGBA8 row
constexpr int F4_XP = 152;              // X row pitch in elements: 76 words, 76 mod 32 = 12 -> conflict-free LDS.128 by row
constexpr int F4_HP = 136;              // [L x64 | H x64] row pitch: 68 words, 68 mod 32 = 4 -> conflict-free STS.128 by row
constexpr int F4_QUADS = F4_XW / 4;     // pixel quads per staged row
constexpr size_t F4_RAW_BYTES = (size_t)FS_STAGES * FS_ROWS * F4_RAW_ROW;
constexpr size_t F4_X_BYTES = (size_t)F4_CH * FS_ROWS * F4_XP * 2;
constexpr size_t F4_HB_BYTES = (size_t)F4_CH * FS_ROWS * F4_HP * 2;
constexpr size_t F4_SMEM = F4_RAW_BYTES + F4_X_BYTES + F4_HB_BYTES + 64;

// ---- colour transform on two pixels at a time: every word holds one component of two pixels as 16-bit lanes.
// Components are 0..255, so sums of two or three stay inside their lane; differences are formed with a bias.

// (a - b) per lane as two's complement int16 lanes, for lanes 0..511 / 0..255; *lt gets 1 in every lane where a < b
__device__ __forceinline__ uint32_t sub2_small(uint32_t a, uint32_t b, uint32_t* lt)
{
	const uint32_t d = a + 0x02000200u - b;                 // 512 + a - b per lane: 257 .. 1023, bit 9 set <=> a >= b
	const uint32_t l = (~d >> 9) & 0x00010001u;
	*lt = l;
	return (d & 0x01FF01FFu) | (l * 0xFE00u);               // a >= b: a - b; a < b: 0xFE00 | (512 + a - b) = a - b mod 2^16
}

// format.c:94-119 (YCoCg / YCoCg_Q), :123-132 (subtract G) for two pixels; identical to color_forward():
//   t = b + (r - b) / 2 with C's truncating division is (r + b + (r < b)) >> 1, and y = t + (g - t) / 2 likewise
__device__ __forceinline__ void color_forward2(int color, uint32_t r, uint32_t g, uint32_t b, uint32_t& p0, uint32_t& p1,
                                               uint32_t& p2)
{
	if (color == AKOD_COL_YCOCG || color == AKOD_COL_YCOCG_Q)
	{
		uint32_t lt_rb, lt_gt;
		const uint32_t co = sub2_small(r, b, &lt_rb);
		const uint32_t t = ((r + b + lt_rb) >> 1) & 0x7FFF7FFFu;
		const uint32_t cg = sub2_small(g, t, &lt_gt);
		const uint32_t y = ((g + t + lt_gt) >> 1) & 0x7FFF7FFFu;
		p0 = (color == AKOD_COL_YCOCG) ? y : (y << 1);
		p1 = co;
		p2 = cg;
	}
	else if (color == AKOD_COL_SUBTRACT_G)
	{
		uint32_t lt;
		p0 = g;
		p1 = sub2_small(r, g, &lt);
		p2 = sub2_small(b, g, &lt);
	}
	else
	{
		p0 = r;
		p1 = g;
		p2 = b;
	}
}

// two RGBA8 pixels -> (r, g, b, a) as 16-bit lanes (lane 0 = the first pixel)
__device__ __forceinline__ void unpack2(uint32_t pa, uint32_t pb, bool discard, uint32_t& r, uint32_t& g, uint32_t& b,
                                        uint32_t& a)
{
	r = __byte_perm(pa, pb, 0x0400) & 0x00FF00FFu; // bytes: a.0, -, b.0, -
	g = __byte_perm(pa, pb, 0x0501) & 0x00FF00FFu;
	b = __byte_perm(pa, pb, 0x0602) & 0x00FF00FFu;
	a = __byte_perm(pa, pb, 0x0703) & 0x00FF00FFu;
	if (discard) // format.c:33-51: colour of invisible pixels is dropped
	{
		const uint32_t keep = ((a & 0xFFFFu) ? 0x0000FFFFu : 0u) | ((a >> 16) ? 0xFFFF0000u : 0u);
		r &= keep;
		g &= keep;
		b &= keep;
	}
}

struct Strip4Params
{
	LiftParams p;            // p.in is unused; channels == 4
	const uint8_t* rgba;     // interleaved RGBA8, image 0
	uint64_t rgba_is;        // bytes between the images of a batch
	uint32_t rgba_rs;        // bytes between rows
	int color, discard;
	uint32_t split;          // coefficient rows per CTA (blockIdx.y)
};

#ifndef F4_CTAS
#define F4_CTAS 3
#endif
template <int WL, int MODE>
__global__ void __launch_bounds__(F4_THREADS, F4_CTAS) k_lift_strip4(const Strip4Params sp)
{
	constexpr bool PLAIN = MODE == FS_PLAIN, GATE = MODE == FS_GATE;
	constexpr int LAT = StripGeom<WL>::LAT;
	const LiftParams& p = sp.p;

	extern __shared__ __align__(128) uint8_t f4_smem[];
	uint8_t* const RAW = f4_smem;                                                       // [stage][row][F4_RAW_ROW]
	int16_t* const X = reinterpret_cast<int16_t*>(f4_smem + F4_RAW_BYTES);              // [ch][row][F4_XP]
	int16_t* const HB = reinterpret_cast<int16_t*>(f4_smem + F4_RAW_BYTES + F4_X_BYTES); // [ch][row][F4_HP]
	uint64_t* const bars = reinterpret_cast<uint64_t*>(f4_smem + F4_RAW_BYTES + F4_X_BYTES + F4_HB_BYTES);

	const int tid = threadIdx.x;
	const int chn = tid >> 6, t64 = tid & 63;
	const uint32_t img = blockIdx.z;
	const int tw = (int)p.tw, th = (int)p.th, cw = (int)p.cw;
	const int c0 = blockIdx.x * F4_TW;
	const int i_begin = blockIdx.y * (int)sp.split;
	const int i_end = min(i_begin + (int)sp.split, th);
	const uint8_t* __restrict__ in = sp.rgba + sp.rgba_is * img;

	const uint32_t band = p.tw * p.th; // < 2^31 elements (host-checked)
	int16_t* __restrict__ ll = p.ll + p.ll_is * img + p.ll_ps * chn;
	int16_t* __restrict__ out_c = p.stream + p.stream_is * img + p.off_c[chn];
	StripQuant sq;
	sq.q = p.q[chn];
	sq.g = p.g[chn];
	sq.mul = p.qmul[chn];
	sq.shift = p.qshift[chn];
	if (blockIdx.x == 0 && blockIdx.y == 0 && t64 == 0)
		out_c[-1] = (int16_t)sq.q; // akoLiftHead

	// ---- loader geometry: staged sample x of a row is image column xs0 + x; [xa, xb) is inside the row
	const int xs0 = 2 * c0 - 8;
	const int xa = (xs0 < 0) ? 8 : 0;
	const int xb = min(F4_XW, cw - xs0);
	const int last_row = (int)p.ch - 1;
	const uint32_t row_bytes = (uint32_t)(xb - xa) * 4;
	const bool edge_strip = (xa != 0) || (xb != F4_XW);
	const int x_right = cw - 2 - xs0; // staged position of the last even column (the CLAMP source on the right)

	auto issue = [&](int js, int buf) {
		if (tid < 32)
		{
			if (tid == 0)
				mbar_expect_tx(&bars[buf], row_bytes * FS_ROWS);
			__syncwarp();
			if (tid < FS_ROWS)
			{
				const int j = min(max(js + (tid >> 1), 0), th - 1);
				const int y = min(2 * j + (tid & 1), last_row);
				bulk_g2s(RAW + (buf * FS_ROWS + tid) * F4_RAW_ROW + xa * 4, in + (uint64_t)y * sp.rgba_rs + (uint64_t)(xs0 + xa) * 4,
				         row_bytes, &bars[buf]);
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

	// ---- V-pass thread geometry (per channel)
	const bool right_half = t64 >= F4_TW / 2;                               // H-pass highpass half -> B, D
	const int vcol = c0 + 2 * (right_half ? t64 - F4_TW / 2 : t64);         // coefficient column of the pair
	const bool vvalid = vcol < tw;
	int16_t* const out_hi = (right_half ? out_c + 2 * (uint64_t)band : out_c) + vcol;
	int16_t* const out_lo = right_half ? (out_c + band + vcol) : (ll + vcol);
	const uint32_t hi_rs = (uint32_t)tw, lo_rs = right_half ? (uint32_t)tw : p.ll_rs;
	const bool odd_offset = ((p.off_c[chn] | p.tw) & 1) != 0;
	const bool vsingle = vcol + 1 >= tw;
	StripV<WL> vs;
	vs.init();
	const int color = sp.color;
	const bool discard = sp.discard != 0;

	const int j_first = i_begin - LAT;
	const int j_last = i_end + LAT; // exclusive
	if (j_first < j_last)
		issue(j_first, 0);

	int buf = 0;
	int16_t* const HBs = HB + chn * (FS_ROWS * F4_HP);
	uint32_t phase = 0;
	for (int js = j_first; js < j_last; js += FS_STEP)
	{
		// the other raw buffer was last read by the conversion of the previous step, which every thread has left
		if (js + FS_STEP < j_last)
			issue(js + FS_STEP, buf ^ 1);
		mbar_wait(&bars[buf], phase);

		// ---------------- conversion: RGBA8 -> the four planes' staged rows (and the CLAMP columns). 16 rows of 144
		// pixels are nine pixels per thread: two quads (one LDS.128 and four STS.64 each) and one single pixel, so
		// that every warp reaches the barrier with the same amount of work done.
		{
			const uint8_t* raw = RAW + buf * (FS_ROWS * F4_RAW_ROW);
			auto source = [&](int r, int x) -> const uint8_t* {
				// CLAMP: every sample outside the row is the first / last even sample
				const int xx = (!edge_strip || (x >= xa && x < xb)) ? x : ((x < xa) ? xa : x_right);
				return raw + r * F4_RAW_ROW + xx * 4;
			};
#pragma unroll
			for (int it = 0; it < 2; it++)
			{
				const int task = tid + it * F4_THREADS; // quads 0 .. 511
				const int r = task / F4_QUADS, q = task - r * F4_QUADS;
				const int x = 4 * q;
				uint4 t;
				if (!edge_strip || (x >= xa && x < xb))
					t = lds128(raw + r * F4_RAW_ROW + x * 4);
				else
				{
					const uint32_t e = *reinterpret_cast<const uint32_t*>(source(r, x));
					t = make_uint4(e, e, e, e);
				}
				uint32_t r0, g0, b0, a0, r1, g1, b1, a1, y0, u0, v0, y1, u1, v1;
				unpack2(t.x, t.y, discard, r0, g0, b0, a0);
				unpack2(t.z, t.w, discard, r1, g1, b1, a1);
				color_forward2(color, r0, g0, b0, y0, u0, v0);
				color_forward2(color, r1, g1, b1, y1, u1, v1);
				int16_t* dst = X + r * F4_XP + x;
				*reinterpret_cast<uint2*>(dst) = make_uint2(y0, y1);
				*reinterpret_cast<uint2*>(dst + FS_ROWS * F4_XP) = make_uint2(u0, u1);
				*reinterpret_cast<uint2*>(dst + 2 * FS_ROWS * F4_XP) = make_uint2(v0, v1);
				*reinterpret_cast<uint2*>(dst + 3 * FS_ROWS * F4_XP) = make_uint2(a0, a1);
			}
			{
				// quads 512 .. 575 one pixel per thread
				const int task = 2 * F4_THREADS + (tid >> 2);
				const int r = task / F4_QUADS, q = task - r * F4_QUADS;
				const int x = 4 * q + (tid & 3);
				const uint32_t e = *reinterpret_cast<const uint32_t*>(source(r, x));
				uint32_t r0, g0, b0, a0, y0, u0, v0;
				unpack2(e, e, discard, r0, g0, b0, a0);
				color_forward2(color, r0, g0, b0, y0, u0, v0);
				int16_t* dst = X + r * F4_XP + x;
				dst[0] = (int16_t)y0;
				dst[FS_ROWS * F4_XP] = (int16_t)u0;
				dst[2 * FS_ROWS * F4_XP] = (int16_t)v0;
				dst[3 * FS_ROWS * F4_XP] = (int16_t)a0;
			}
		}
		__syncthreads();

		// ---------------- H pass: thread = (row, chunk of 16 coefficient pairs) of its channel
		{
			const int r = t64 & 15, chunk = t64 >> 4;
			const int a = chunk * 16;
			if (c0 + a < tw)
			{
				uint32_t w[24];
				const int16_t* src = X + (chn * FS_ROWS + r) * F4_XP + 2 * a;
#pragma unroll
				for (int k = 0; k < 6; k++)
				{
					const uint4 t = lds128(src + 8 * k);
					w[4 * k] = t.x;
					w[4 * k + 1] = t.y;
					w[4 * k + 2] = t.z;
					w[4 * k + 3] = t.w;
				}
				uint32_t lw[8], hw[8];
				strip_hpass<WL>(w, c0 + a == 0, tw - (c0 + a), lw, hw);
				uint4* dl = reinterpret_cast<uint4*>(&HBs[r * F4_HP + a]);
				uint4* dh = reinterpret_cast<uint4*>(&HBs[r * F4_HP + F4_TW + a]);
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
			const uint32_t* col = reinterpret_cast<const uint32_t*>(HBs) + t64;
			const int i0 = js - LAT; // output row of this step's first input row
			const bool interior = (i0 >= i_begin) && (i0 + FS_STEP <= i_end) && (js > LAT) && (js + FS_STEP <= th);
			int16_t* const row_hi = out_hi + (int64_t)i0 * (int64_t)hi_rs;
			int16_t* const row_lo = out_lo + (int64_t)i0 * (int64_t)lo_rs;

			auto vstep = [&](auto edge_tag, auto odd_tag) {
				constexpr bool EDGE = decltype(edge_tag)::value;
				constexpr bool ODD = decltype(odd_tag)::value;
#pragma unroll
				for (int k = 0; k < FS_STEP; k++)
				{
					const uint32_t we = col[(2 * k) * (F4_HP / 2)], wo = col[(2 * k + 1) * (F4_HP / 2)];
					int la, lb, ha, hb2;
					vs.template row<EDGE>(we, wo, js + k, th, la, lb, ha, hb2);
					if (!EDGE || (uint32_t)(i0 + k - i_begin) < (uint32_t)(i_end - i_begin))
					{
						int16_t* dh = row_hi + (uint32_t)k * hi_rs;
						int16_t* dl = row_lo + (uint32_t)k * lo_rs;
						uint32_t whi, wlo;
						if (PLAIN)
							whi = pair_hi((uint32_t)ha, (uint32_t)hb2);
						else
							whi = pack2(strip_quant<GATE>(ha, sq), strip_quant<GATE>(hb2, sq));
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
		// No barrier here. Two CTA-wide barriers per step order everything: a thread reaches the barrier after the next
		// conversion only when it has left this V pass, so the next H pass (behind that barrier) may overwrite HB; the
		// next conversion rewrites X, which every H pass has left (barrier above); the raw buffer refilled at the top
		// of the next step was last read by the conversion of the step before this one.
		buf ^= 1;
		if (buf == 0)
			phase ^= 1u;
	}
}

// host-side eligibility of level 0 for the fused kernel (p as akod_lift builds it, without p.in)
static inline bool lift_strip4_eligible(const LiftParams& p, const uint8_t* rgba, uint64_t rgba_is, uint64_t rgba_rs)
{
	return p.channels == 4 && p.wrap == AKOD_WRAP_CLAMP && (p.cw % 4) == 0 && p.cw >= 64 && p.th >= 8 &&
	       ((uintptr_t)rgba % 16) == 0 && (rgba_is % 16) == 0 && (rgba_rs % 16) == 0 && rgba_rs < ((uint64_t)1 << 32) &&
	       (p.ll_rs % 2) == 0 && (p.ll_ps % 2) == 0 && (p.ll_is % 2) == 0 && ((uintptr_t)p.ll % 4) == 0 &&
	       (p.stream_is % 2) == 0 && ((uintptr_t)p.stream % 4) == 0 && (uint64_t)p.cw * p.ch < ((uint64_t)1 << 31);
}
