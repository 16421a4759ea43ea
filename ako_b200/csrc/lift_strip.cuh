// lift_strip.cuh -- the fast forward-lifting kernel ("strip marching"), used for the large, aligned,
// CLAMP-wrapped levels that carry nearly all the bytes. Same arithmetic as k_lift_level (lift.cuh), which
// stays the general kernel for every other case (other wrap modes, odd widths, tiny levels).
//
// Why it looks like this. At 60 % of the HBM roofline a B200 SM has about 16 issue slots per int16 sample
// for a whole 2-D level, so the lifting taps must live in registers, not be re-read from shared memory:
//   * a CTA owns a strip of 128 coefficient columns (256 samples + 8 halo each side) and MARCHES down it,
//     16 input rows (8 coefficient rows) per step;
//   * H pass: one thread per (row, 16-pair chunk): six 128-bit shared loads, the 19 highpass + 16 lowpass
//     values computed from registers, four 128-bit shared stores. Only the 3 halo highpasses are redundant;
//   * V pass: one thread per pair of adjacent columns for the WHOLE strip height: its sliding window of
//     even rows and highpass values stays in registers from step to step, so nothing is recomputed and no
//     vertical halo is ever re-loaded. Results go straight to global memory;
//   * the next step's 16 rows are prefetched into registers (128-bit loads) while the current step computes.
// Boundary rules (CLAMP): the loader clamps row/column indices, which gives E(-1)=E(0), E(t)=E(t+1)=E(t-1)
// and the duplicated last row of odd heights; the highpass overrides H(-1)=H(-2)=H(0) and H(t)=H(t-1) are
// applied where those values are produced (see wavelet-dd137.c:76-77, :110-111, :122, :151-152, :163).
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
constexpr int FS_VEC_PER_ROW = FS_XW / 8;                    // 34 128-bit vectors per staged row
constexpr int FS_VECS = FS_ROWS * FS_VEC_PER_ROW;            // 544 per step
constexpr int FS_PREFETCH = (FS_VECS + FS_THREADS - 1) / FS_THREADS; // 5 per thread

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

__device__ __forceinline__ int sx16(int v) // narrow to int16 by wrap, as every store into an int16_t does
{
	// (PTX prmt's sign-replicate selector is NOT honoured by ptxas for sm_100a -- it becomes a byte copy --
	// so this stays a plain cast, which compiles to one sign-extension instruction.)
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

// the divisions are written as C divisions by constants (toward zero); the compiler turns them into 2-3 ops
template <int WL>
__device__ __forceinline__ int strip_hp(int o, int e, int l1, int p1, int p2)
{
	if (WL == AKOD_HAAR)
		return sx16(o - e);
	if (WL == AKOD_CDF53)
		return sx16(o - (e + p1) / 2);
	return sx16(o + (l1 + p2 - 9 * (e + p1)) / 16);
}

template <int WL>
__device__ __forceinline__ int strip_lp(int e, int l2, int l1, int h, int p1)
{
	if (WL == AKOD_HAAR)
		return e;
	if (WL == AKOD_CDF53)
		return e + (l1 + h) / 4;
	return e + (-l2 - p1 + 9 * (l1 + h)) / 32;
}

__device__ __forceinline__ uint32_t pack2(int lo, int hi)
{
	return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410);
}

// gate + quantise of a value already narrowed to int16 (lifting.c:163)
__device__ __forceinline__ int strip_quant(int v, int q, int g, uint32_t magic)
{
	const int a = abs(v);
	int d = (q > 1) ? (int)__umulhi((uint32_t)a, magic) : a;
	d = (v < 0) ? -d : d;
	return (a > g) ? d : 0;
}

// PLAIN: q == 1 and gate == 0 on every channel (the quantise step is the identity).
template <int WL, bool PLAIN>
__global__ void __launch_bounds__(FS_THREADS, 6) k_lift_strip(const StripParams sp)
{
	using SG = StripGeom<WL>;
	constexpr int LAT = SG::LAT;
	const LiftParams& p = sp.p;

	__shared__ __align__(16) int16_t X[FS_ROWS * FS_XP];
	__shared__ __align__(16) int16_t HB[FS_ROWS * FS_HP];

	const int tid = threadIdx.x;
	const uint32_t img = blockIdx.z / p.channels, chn = blockIdx.z - img * p.channels;
	const int tw = (int)p.tw, th = (int)p.th;
	const int c0 = blockIdx.x * FS_TW;
	const int i_begin = blockIdx.y * (int)sp.split;
	const int i_end = min(i_begin + (int)sp.split, th);
	const int16_t* __restrict__ in = p.in + p.in_is * img + p.in_ps * chn;

	const uint32_t band = p.tw * p.th; // < 2^31 elements (host-checked)
	int16_t* __restrict__ ll = p.ll + p.ll_is * img + p.ll_ps * chn;
	int16_t* __restrict__ out_c = p.stream + p.stream_is * img + p.off_c[chn];
	const int q = p.q[chn], g = p.g[chn];
	const uint32_t magic = p.qmagic[chn];
	if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0)
		out_c[-1] = (int16_t)q; // akoLiftHead

	// ---- loader. Per thread and slot the staged row, the column offset and the clamp case never change
	int pre_row[FS_PREFETCH], pre_x[FS_PREFETCH]; // pre_x < 0: clamp (-1: left edge, -2: right edge, -3: no slot)
#pragma unroll
	for (int k = 0; k < FS_PREFETCH; k++)
	{
		const int i = tid + FS_THREADS * k;
		const int r = i / FS_VEC_PER_ROW, v = i - r * FS_VEC_PER_ROW;
		const int x = 2 * c0 - 8 + 8 * v;
		pre_row[k] = r;
		pre_x[k] = (i >= FS_VECS) ? -3 : (x < 0) ? -1 : (x + 8 > (int)p.cw) ? -2 : x;
	}
	const int in_rs = (int)p.in_rs, last_row = (int)p.ch - 1, cw2 = (int)p.cw - 2;
	uint4 pre[FS_PREFETCH];
	auto prefetch = [&](int js) {
#pragma unroll
		for (int k = 0; k < FS_PREFETCH; k++)
		{
			if (pre_x[k] != -3)
			{
				const int j = min(max(js + (pre_row[k] >> 1), 0), th - 1);
				const int y = min(2 * j + (pre_row[k] & 1), last_row);
				const int16_t* row = in + (uint32_t)(y * in_rs); // planes are < 2^31 elements (host-checked)
				if (pre_x[k] >= 0)
					pre[k] = __ldg(reinterpret_cast<const uint4*>(row + pre_x[k]));
				else
				{
					// CLAMP: every even sample outside the row is the first / last even sample
					const uint32_t e = (uint16_t)__ldg(row + (pre_x[k] == -1 ? 0 : cw2));
					const uint32_t w = e * 0x10001u;
					pre[k] = make_uint4(w, w, w, w);
				}
			}
		}
	};
	auto commit = [&]() {
#pragma unroll
		for (int k = 0; k < FS_PREFETCH; k++)
			if (pre_x[k] != -3)
			{
				const int i = tid + FS_THREADS * k;
				const int v = i - pre_row[k] * FS_VEC_PER_ROW;
				*reinterpret_cast<uint4*>(&X[pre_row[k] * FS_XP + 8 * v]) = pre[k];
			}
	};

	// ---- V-pass state: two adjacent columns per thread
	const bool right_half = tid >= FS_TW / 2;                               // H-pass highpass half -> B, D
	const int vcol = c0 + 2 * (right_half ? tid - FS_TW / 2 : tid);         // coefficient column of the pair
	const bool vvalid = vcol < tw;
	int e1[2] = {0, 0}, e2[2] = {0, 0}, e3[2] = {0, 0}, o1[2] = {0, 0}, o2[2] = {0, 0};
	int ha[2] = {0, 0}, hb[2] = {0, 0}, hc[2] = {0, 0};
	int16_t* out_hi = (right_half ? out_c + 2 * (uint64_t)band : out_c) + vcol; // V-high: D or C
	int16_t* out_lo = right_half ? (out_c + band + vcol) : (ll + vcol);          // V-low: B, or the next level's input

	// The C/B/D subbands of a channel start at an odd or even int16 offset of the stream (a 2-byte lift head
	// precedes each channel's block, so the parity alternates from channel to channel): pairs are stored with
	// one 32-bit store when aligned, two 16-bit stores otherwise. Uniform per CTA; resolved outside the row loop.
	const bool odd_offset = (p.off_c[chn] & 1) != 0;

	const int j_first = i_begin - LAT;
	const int j_last = i_end + LAT; // exclusive
	const uint32_t n_out = (uint32_t)(i_end - i_begin);
	prefetch(j_first);

	for (int js = j_first; js < j_last; js += FS_STEP)
	{
		commit();
		__syncthreads();
		if (js + FS_STEP < j_last)
			prefetch(js + FS_STEP);

		// ---------------- H pass: thread = (row, chunk of 16 coefficient pairs)
		{
			const int r = tid & 15, chunk = tid >> 4;
			const int a = chunk * 16;
			if (c0 + a < tw)
			{
				// words [a, a+24) of the staged row hold pairs c = c0 + a - 4 + k
				uint32_t w[24];
				const uint4* src = reinterpret_cast<const uint4*>(&X[r * FS_XP + 2 * a]);
#pragma unroll
				for (int k = 0; k < 6; k++)
				{
					const uint4 t = src[k];
					w[4 * k] = t.x;
					w[4 * k + 1] = t.y;
					w[4 * k + 2] = t.z;
					w[4 * k + 3] = t.w;
				}
				int E[24], H[24];
#pragma unroll
				for (int k = 1; k < 23; k++)
					E[k] = lo16(w[k]);
#pragma unroll
				for (int k = 2; k <= 20; k++)
				{
					if (WL == AKOD_DD137)
						H[k] = strip_hp<WL>(hi16(w[k]), E[k], E[k - 1], E[k + 1], E[k + 2]);
					else
						H[k] = strip_hp<WL>(hi16(w[k]), E[k], 0, E[k + 1], 0);
				}
				if (WL != AKOD_HAAR)
				{
					if (c0 + a == 0) // H(-1) = H(-2) = H(0)
						H[2] = H[3] = H[4];
					const int rem = tw - (c0 + a); // H(t) = H(t-1); t is 4, 8, 12 or 16 columns into an edge chunk
					if (rem == 4)
						H[8] = H[7];
					if (rem == 8)
						H[12] = H[11];
					if (rem == 12)
						H[16] = H[15];
					if (rem == 16)
						H[20] = H[19];
				}
				uint32_t lw[8], hw[8];
#pragma unroll
				for (int k = 0; k < 8; k++)
				{
					const int k0 = 4 + 2 * k, k1 = k0 + 1;
					const int l0 = strip_lp<WL>(E[k0], H[k0 - 2], H[k0 - 1], H[k0], H[k0 + 1]);
					const int l1 = strip_lp<WL>(E[k1], H[k1 - 2], H[k1 - 1], H[k1], H[k1 + 1]);
					lw[k] = pack2(l0, l1);
					hw[k] = pack2(H[k0], H[k1]);
				}
				uint4* dl = reinterpret_cast<uint4*>(&HB[r * FS_HP + a]);
				uint4* dh = reinterpret_cast<uint4*>(&HB[r * FS_HP + FS_TW + a]);
				dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
				dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
				dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
				dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
			}
		}
		__syncthreads();

		// ---------------- V pass: marching, state in registers
		auto vpass = [&](auto odd_tag) {
			constexpr bool ODD = decltype(odd_tag)::value;
			auto store2 = [&](int16_t* dst, int a, int b) {
				if (ODD)
				{
					dst[0] = (int16_t)a;
					dst[1] = (int16_t)b;
				}
				else
					*reinterpret_cast<uint32_t*>(dst) = pack2(a, b);
			};
			const uint32_t* col = reinterpret_cast<const uint32_t*>(HB) + tid;
			// row offsets of this step's first output row (may be before i_begin: then nothing is stored)
			const int i0 = js - LAT;
			int16_t* row_hi = out_hi + (int64_t)i0 * tw;
			int16_t* row_lo = out_lo + (int64_t)i0 * (right_half ? tw : (int)p.ll_rs);
			const int lo_step = right_half ? tw : (int)p.ll_rs;
#pragma unroll
			for (int k = 0; k < FS_STEP; k++)
			{
				const int j = js + k;
				const uint32_t we = col[(2 * k) * (FS_HP / 2)], wo = col[(2 * k + 1) * (FS_HP / 2)];
				const int ej[2] = {lo16(we), hi16(we)};
				const int oj[2] = {lo16(wo), hi16(wo)};
				int lo[2], hi[2];
#pragma unroll
				for (int s = 0; s < 2; s++)
				{
					if (WL == AKOD_DD137)
					{
						int h = strip_hp<WL>(o2[s], e2[s], e3[s], e1[s], ej[s]); // H(j-2)
						if (j - 2 == 0)
							ha[s] = hb[s] = h; // H(-1) = H(-2) = H(0)
						if (j - 2 >= th)
							h = ha[s]; // H(t) = H(t-1)
						lo[s] = strip_lp<WL>(e3[s], hc[s], hb[s], ha[s], h); // L(j-3)
						hi[s] = ha[s];                                       // H(j-3)
						hc[s] = hb[s];
						hb[s] = ha[s];
						ha[s] = h;
						e3[s] = e2[s];
						e2[s] = e1[s];
						e1[s] = ej[s];
						o2[s] = o1[s];
						o1[s] = oj[s];
					}
					else if (WL == AKOD_CDF53)
					{
						const int h = strip_hp<WL>(o1[s], e1[s], 0, ej[s], 0); // H(j-1)
						if (j - 1 == 0)
							ha[s] = h; // H(-1) = H(0)
						lo[s] = strip_lp<WL>(e1[s], 0, ha[s], h, 0); // L(j-1)
						hi[s] = h;
						ha[s] = h;
						e1[s] = ej[s];
						o1[s] = oj[s];
					}
					else
					{
						lo[s] = ej[s];
						hi[s] = strip_hp<WL>(oj[s], ej[s], 0, 0, 0);
					}
				}
				if ((uint32_t)(i0 + k - i_begin) < n_out)
				{
					if (!PLAIN)
					{
						hi[0] = strip_quant(hi[0], q, g, magic);
						hi[1] = strip_quant(hi[1], q, g, magic);
					}
					store2(row_hi, hi[0], hi[1]);
					if (right_half)
					{
						if (!PLAIN)
						{
							lo[0] = strip_quant(sx16(lo[0]), q, g, magic);
							lo[1] = strip_quant(sx16(lo[1]), q, g, magic);
						}
						store2(row_lo, lo[0], lo[1]);
					}
					else
						*reinterpret_cast<uint32_t*>(row_lo) = pack2(lo[0], lo[1]);
				}
				row_hi += tw;
				row_lo += lo_step;
			}
		};
		if (vvalid)
		{
			if (odd_offset)
				vpass(std::true_type{});
			else
				vpass(std::false_type{});
		}
		__syncthreads(); // HB and X are rewritten by the next step
	}
}

// host-side eligibility of a level for the strip kernel
static inline bool lift_strip_eligible(const LiftParams& p)
{
	return p.wrap == AKOD_WRAP_CLAMP && (p.cw % 8) == 0 && p.cw >= 64 && p.th >= 8 && (p.in_rs % 8) == 0 &&
	       (p.in_ps % 8) == 0 && (p.in_is % 8) == 0 && ((uintptr_t)p.in % 16) == 0 && (p.ll_rs % 2) == 0 &&
	       (p.ll_ps % 2) == 0 && (p.ll_is % 2) == 0 && ((uintptr_t)p.ll % 4) == 0 && (p.stream_is % 2) == 0 &&
	       ((uintptr_t)p.stream % 4) == 0 && (uint64_t)p.cw * p.ch < ((uint64_t)1 << 31);
}
