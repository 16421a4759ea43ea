// lift_small.cuh -- the tail of the pyramid in one launch. Once a level's plane fits in shared memory
// (cw*ch <= 32768 samples) a single CTA per (image, channel) runs ALL remaining levels there: H pass, V pass,
// quantise + gate, subband stores, and finally the lowpass section of the stream (lifting.c:171-292); the
// inverse kernel does the same walk coarsest-first (lifting.c:86-148, :295-304). This replaces 6-8 dependent
// launches per direction whose cost was latency, not bytes: e.g. for 8192x8192 the levels below 128x128 took
// ~100 us of a 410 us forward transform as separate kernels.
//
// The arithmetic is the plain statement of the transform (oracle/ako_oracle.c restates the same thing on the
// CPU): for every coefficient the taps go through the wrap mode's index map, so all four wrap modes, odd sizes
// (plus-one rule) and the DD137 -> CDF53 fallback of tiny levels are handled here.
#pragma once

#include "lift.cuh"

constexpr int SM_THREADS = 1024;      // planes of up to 32 768 samples
constexpr int SM_THREADS_TILE = 256;  // planes of up to 8 192 samples (tile batches): several CTAs share an SM
constexpr int SM_MAX_LEVELS = 16;
constexpr uint32_t SM_MAX_SAMPLES = 32768;      // plane size at which the small kernels take over
constexpr uint32_t SM_CAP = 34816;              // elements per shared buffer (two buffers, 136 KiB)

struct SmallLevelQ
{
	int16_t qy, qc, gy, gc; // channel 0 / every other channel (lifting.c:202-211)
};

struct SmallParams
{
	// forward: planes of the first small level (input). inverse: planes of the same level (output)
	int16_t* planes;
	uint32_t planes_rs;
	uint64_t planes_ps, planes_is;
	int16_t* stream;
	uint64_t stream_is;
	uint32_t cw0, ch0;   // dimensions of the first (finest) small level's plane
	uint32_t levels;     // how many levels the kernel runs (all that remain)
	uint32_t channels;
	uint32_t cap;        // elements per shared buffer (small_capacity): a 64 x 64 tile does not ask for 136 KiB
	int32_t wrap, wavelet; // wavelet = the settings' wavelet; the per-level fallback is applied here
	SmallLevelQ lq[SM_MAX_LEVELS]; // [0] = finest small level
};

__host__ __device__ inline uint32_t sm_half(uint32_t v) // akoDividePlusOneRule
{
	return (v + 1) / 2;
}

__device__ __forceinline__ int sm_level_wavelet(int wavelet, int tw, int th) // lifting.c:49, :58, :67
{
	if (wavelet == AKOD_HAAR)
		return AKOD_HAAR;
	if (wavelet == AKOD_CDF53 || tw < 8 || th < 8)
		return AKOD_CDF53;
	return AKOD_DD137;
}

// Walks a rows x cols index space with all NT threads of the CTA, element by element (row-major), without a division
// per element: the small levels have fewer columns than a warp has lanes, so a (warp = row, lane = column)
// mapping would leave most lanes idle.
template <int NT>
struct SmWalk
{
	int row, col, dq, dr, cols;
	__device__ __forceinline__ SmWalk(int cols_) : cols(cols_)
	{
		row = (int)threadIdx.x / cols;
		col = (int)threadIdx.x - row * cols;
		dq = NT / cols;
		dr = NT - dq * cols;
	}
	__device__ __forceinline__ void next()
	{
		col += dr;
		row += dq;
		if (col >= cols)
		{
			col -= cols;
			row++;
		}
	}
};

// tap through the wrap mode's index map (period t, element stride 'stride'); zero outside for WRAP_ZERO.
// CLAMP (the default wrap mode, known at compile time in the level functions) is a branch-free index clamp.
template <bool CLAMP>
__device__ __forceinline__ int sm_tap(const int16_t* a, int stride, int wrap, int v, int t)
{
	if (CLAMP)
		return (int)a[min(max(v, 0), t - 1) * stride];
	const int m = wrap_map(v, t, wrap);
	return (m < 0) ? 0 : (int)a[m * stride];
}

// offsets (int16 units) of the C subband of every small level for one channel; lifting.c:179-291, misc.c:229-285
__device__ __forceinline__ void sm_offsets(const SmallParams& p, uint32_t chn, uint32_t* off_c, uint32_t* lw, uint32_t* lh)
{
	// lw[s], lh[s] = input plane of small level s; lw[levels], lh[levels] = final lowpass
	lw[0] = p.cw0;
	lh[0] = p.ch0;
	for (uint32_t s = 0; s < p.levels; s++)
	{
		lw[s + 1] = sm_half(lw[s]);
		lh[s + 1] = sm_half(lh[s]);
	}
	uint32_t base = p.channels * lw[p.levels] * lh[p.levels]; // the lowpass section comes first
	for (uint32_t s = p.levels; s-- > 0;)
	{
		const uint32_t block = 1 + 3 * lw[s + 1] * lh[s + 1]; // head + C + B + D
		off_c[s] = base + chn * block + 1;
		base += p.channels * block;
	}
}

// ------------------------------------------------------------------------------------------------
// forward

template <int WL, bool CLAMP, int NT>
__device__ __forceinline__ void sm_forward_level(int16_t* A, int16_t* Bf, int cw, int ch, int tw, int th, int wrap,
                                                 int q, int g, uint32_t magic, int16_t* out_c, int16_t* ll_next)
{
	const int bw = 2 * tw; // row pitch of the H-pass output [L | H]

	// ---- H pass, highpass: Bf[y][tw + c]
	for (SmWalk<NT> it(tw); it.row < ch; it.next())
	{
		const int y = it.row, c = it.col;
		const int16_t* row = A + y * cw;
		{
			const int e = row[2 * c], o = row[min(2 * c + 1, cw - 1)]; // odd width: duplicate last column
			int h;
			if (WL == AKOD_DD137)
			{
				const int l1 = sm_tap<CLAMP>(row, 2, wrap, c - 1, tw), p1 = sm_tap<CLAMP>(row, 2, wrap, c + 1, tw);
				const int p2 = (!CLAMP && wrap == AKOD_WRAP_MIRROR && c >= tw - 2) ? l1 : sm_tap<CLAMP>(row, 2, wrap, c + 2, tw);
				h = hp_forward<WL>(o, e, l1, p1, p2);
			}
			else if (WL == AKOD_CDF53)
				h = hp_forward<WL>(o, e, 0, sm_tap<CLAMP>(row, 2, wrap, c + 1, tw), 0);
			else
				h = hp_forward<WL>(o, e, 0, 0, 0);
			Bf[y * bw + tw + c] = (int16_t)h;
		}
	}
	__syncthreads();
	// ---- H pass, lowpass: Bf[y][c]
	for (SmWalk<NT> it(tw); it.row < ch; it.next())
	{
		const int y = it.row, c = it.col;
		const int16_t* row = A + y * cw;
		const int16_t* hrow = Bf + y * bw + tw;
		{
			const int e = row[2 * c];
			int l;
			if (WL == AKOD_DD137)
			{
				const int l1 = sm_tap<CLAMP>(hrow, 1, wrap, c - 1, tw), p1 = sm_tap<CLAMP>(hrow, 1, wrap, c + 1, tw);
				const int l2 = (!CLAMP && wrap == AKOD_WRAP_MIRROR && c <= 1) ? p1 : sm_tap<CLAMP>(hrow, 1, wrap, c - 2, tw);
				l = lp_forward<WL>(e, l2, l1, hrow[c], p1);
			}
			else if (WL == AKOD_CDF53)
				l = lp_forward<WL>(e, 0, sm_tap<CLAMP>(hrow, 1, wrap, c - 1, tw), hrow[c], 0);
			else
				l = e;
			Bf[y * bw + c] = (int16_t)l;
		}
	}
	__syncthreads();
	// ---- V pass, highpass of every column of [L | H]: HV[r][col] overwrites A (dead now)
	int16_t* HV = A;
	for (SmWalk<NT> it(bw); it.row < th; it.next())
	{
		const int r = it.row, col = it.col;
		{
			const int16_t* colp = Bf + col;
			const int e = colp[(2 * r) * bw], o = colp[min(2 * r + 1, ch - 1) * bw]; // odd height: duplicate last row
			int h;
			if (WL == AKOD_DD137)
			{
				const int l1 = sm_tap<CLAMP>(colp, 2 * bw, wrap, r - 1, th), p1 = sm_tap<CLAMP>(colp, 2 * bw, wrap, r + 1, th);
				const int p2 = (!CLAMP && wrap == AKOD_WRAP_MIRROR && r >= th - 2) ? l1 : sm_tap<CLAMP>(colp, 2 * bw, wrap, r + 2, th);
				h = hp_forward<WL>(o, e, l1, p1, p2);
			}
			else if (WL == AKOD_CDF53)
				h = hp_forward<WL>(o, e, 0, sm_tap<CLAMP>(colp, 2 * bw, wrap, r + 1, th), 0);
			else
				h = hp_forward<WL>(o, e, 0, 0, 0);
			HV[r * bw + col] = (int16_t)h;
		}
	}
	__syncthreads();
	// ---- V pass, lowpass + gate/quantise + stores. LL goes to ll_next (shared), C/B/D to the stream.
	int16_t* out_b = out_c + tw * th;
	int16_t* out_d = out_b + tw * th;
	for (SmWalk<NT> it(bw); it.row < th; it.next())
	{
		const int r = it.row, col = it.col;
		{
			const int16_t* hcol = HV + col;
			const int e = Bf[(2 * r) * bw + col];
			const int hv = hcol[r * bw];
			int l;
			if (WL == AKOD_DD137)
			{
				const int l1 = sm_tap<CLAMP>(hcol, bw, wrap, r - 1, th), p1 = sm_tap<CLAMP>(hcol, bw, wrap, r + 1, th);
				const int l2 = (!CLAMP && wrap == AKOD_WRAP_MIRROR && r <= 1) ? p1 : sm_tap<CLAMP>(hcol, bw, wrap, r - 2, th);
				l = lp_forward<WL>(e, l2, l1, hv, p1);
			}
			else if (WL == AKOD_CDF53)
				l = lp_forward<WL>(e, 0, sm_tap<CLAMP>(hcol, bw, wrap, r - 1, th), hv, 0);
			else
				l = e;
			if (col < tw)
			{
				ll_next[r * tw + col] = (int16_t)l;
				out_c[r * tw + col] = gate_quantize(hv, q, g, magic);
			}
			else
			{
				out_b[r * tw + col - tw] = gate_quantize((int16_t)l, q, g, magic);
				out_d[r * tw + col - tw] = gate_quantize(hv, q, g, magic);
			}
		}
	}
	__syncthreads();
}

template <int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 4 : 1) k_lift_small(const SmallParams p)
{
	extern __shared__ __align__(16) int16_t sm_buf[];
	int16_t* A = sm_buf;
	int16_t* Bf = sm_buf + p.cap;
	__shared__ uint32_t off_c[SM_MAX_LEVELS], lw[SM_MAX_LEVELS + 1], lh[SM_MAX_LEVELS + 1];

	const uint32_t img = blockIdx.x / p.channels, chn = blockIdx.x - img * p.channels;
	if (threadIdx.x == 0)
		sm_offsets(p, chn, off_c, lw, lh);
	const int16_t* in = p.planes + p.planes_is * img + p.planes_ps * chn;
	int16_t* stream = p.stream + p.stream_is * img;
	for (SmWalk<NT> it((int)p.cw0); it.row < (int)p.ch0; it.next())
		A[it.row * (int)p.cw0 + it.col] = __ldg(in + (uint64_t)it.row * p.planes_rs + it.col);
	__syncthreads();

	int16_t* cur = A; // dense lw[s] x lh[s]
	for (uint32_t s = 0; s < p.levels; s++)
	{
		const int cw = (int)lw[s], ch = (int)lh[s], tw = (int)lw[s + 1], th = (int)lh[s + 1];
		const int wl = sm_level_wavelet(p.wavelet, tw, th);
		int q = chn == 0 ? p.lq[s].qy : p.lq[s].qc;
		const int g = chn == 0 ? p.lq[s].gy : p.lq[s].gc;
		q = q < 1 ? 1 : q;
		const uint32_t magic = (q > 1) ? (uint32_t)((((uint64_t)1 << 32) + q - 1) / (uint64_t)q) : 0;
		int16_t* out_c = stream + off_c[s];
		if (threadIdx.x == 0)
			out_c[-1] = (int16_t)q; // akoLiftHead
		int16_t* ll_next = cur + th * 2 * tw; // behind the V-highpass scratch that overwrites 'cur'
		const bool clamp = p.wrap == AKOD_WRAP_CLAMP;
		if (wl == AKOD_DD137 && clamp)
			sm_forward_level<AKOD_DD137, true, NT>(cur, Bf, cw, ch, tw, th, p.wrap, q, g, magic, out_c, ll_next);
		else if (wl == AKOD_DD137)
			sm_forward_level<AKOD_DD137, false, NT>(cur, Bf, cw, ch, tw, th, p.wrap, q, g, magic, out_c, ll_next);
		else if (wl == AKOD_CDF53 && clamp)
			sm_forward_level<AKOD_CDF53, true, NT>(cur, Bf, cw, ch, tw, th, p.wrap, q, g, magic, out_c, ll_next);
		else if (wl == AKOD_CDF53)
			sm_forward_level<AKOD_CDF53, false, NT>(cur, Bf, cw, ch, tw, th, p.wrap, q, g, magic, out_c, ll_next);
		else
			sm_forward_level<AKOD_HAAR, true, NT>(cur, Bf, cw, ch, tw, th, p.wrap, q, g, magic, out_c, ll_next);
		cur = ll_next;
	}
	// lowpass section (lifting.c:280-291)
	const uint32_t lpn = lw[p.levels] * lh[p.levels];
	for (uint32_t i = threadIdx.x; i < lpn; i += NT)
		stream[chn * lpn + i] = cur[i];
}

// ------------------------------------------------------------------------------------------------
// inverse

template <int WL, bool CLAMP, int NT>
__device__ __forceinline__ void sm_inverse_level(int16_t* R, int16_t* T, int hw, int hh, int tw, int th, int wrap, int q,
                                                 const int16_t* __restrict__ in_c)
{
	// R = [LL hw*hh | C | B | D] (the three subbands staged here, inverse-quantised); result (tw x th, dense)
	// is written back to R[0..). T = 2hh rows x [left hw | right hw].
	const int band = hw * hh, bw = 2 * hw;
	int16_t* LL = R;
	int16_t* HC = R + band; // C, then B, then D

	// lifting.c:30-40: coefficient * q narrowed to int16, skipped when q <= 1
	for (int i = threadIdx.x; i < 3 * band; i += NT)
	{
		const int v = __ldg(in_c + i);
		HC[i] = (int16_t)((q > 1) ? v * q : v);
	}
	__syncthreads();
	// ---- V pass, even rows: T[2r][col]
	for (SmWalk<NT> it(bw); it.row < hh; it.next())
	{
		const int r = it.row, col = it.col;
		{
			const bool left = col < hw;
			const int c = left ? col : col - hw;
			const int lp = left ? LL[r * hw + c] : HC[band + r * hw + c];              // LL or B
			const int16_t* hcol = (left ? HC : HC + 2 * band) + c;                      // C or D column
			const int h0 = hcol[r * hw];
			int e;
			if (WL == AKOD_DD137)
			{
				const int l1 = sm_tap<CLAMP>(hcol, hw, wrap, r - 1, hh), p1 = sm_tap<CLAMP>(hcol, hw, wrap, r + 1, hh);
				const int l2 = (!CLAMP && wrap == AKOD_WRAP_MIRROR && r <= 1) ? p1 : sm_tap<CLAMP>(hcol, hw, wrap, r - 2, hh);
				e = even_inverse<WL>(lp, l2, l1, h0, p1);
			}
			else if (WL == AKOD_CDF53)
				e = even_inverse<WL>(lp, 0, sm_tap<CLAMP>(hcol, hw, wrap, r - 1, hh), h0, 0);
			else
				e = lp;
			T[(2 * r) * bw + col] = (int16_t)e;
		}
	}
	__syncthreads();
	// ---- V pass, odd rows: T[2r+1][col]
	for (SmWalk<NT> it(bw); it.row < hh; it.next())
	{
		const int r = it.row, col = it.col;
		{
			const bool left = col < hw;
			const int c = left ? col : col - hw;
			const int h0 = (left ? HC : HC + 2 * band)[r * hw + c];
			const int16_t* ecol = T + col;
			const int e0 = ecol[(2 * r) * bw];
			int o;
			if (WL == AKOD_DD137)
			{
				const int l1 = sm_tap<CLAMP>(ecol, 2 * bw, wrap, r - 1, hh), p1 = sm_tap<CLAMP>(ecol, 2 * bw, wrap, r + 1, hh);
				const int p2 = (!CLAMP && wrap == AKOD_WRAP_MIRROR && r >= hh - 2) ? l1 : sm_tap<CLAMP>(ecol, 2 * bw, wrap, r + 2, hh);
				o = odd_inverse<WL>(h0, e0, l1, p1, p2);
			}
			else if (WL == AKOD_CDF53)
				o = odd_inverse<WL>(h0, e0, 0, sm_tap<CLAMP>(ecol, 2 * bw, wrap, r + 1, hh), 0);
			else
				o = odd_inverse<WL>(h0, e0, 0, 0, 0);
			T[(2 * r + 1) * bw + col] = (int16_t)o;
		}
	}
	__syncthreads();
	// ---- H pass, even samples of the th real rows: R[y][2c]  (LL and the staged subbands are dead now)
	const int hwt = (tw + 1) / 2; // == hw
	for (SmWalk<NT> it(hwt); it.row < th; it.next())
	{
		const int y = it.row, c = it.col;
		const int16_t* lrow = T + y * bw;
		const int16_t* hrow = lrow + hw;
		{
			int e;
			if (WL == AKOD_DD137)
			{
				const int l1 = sm_tap<CLAMP>(hrow, 1, wrap, c - 1, hw), p1 = sm_tap<CLAMP>(hrow, 1, wrap, c + 1, hw);
				const int l2 = (!CLAMP && wrap == AKOD_WRAP_MIRROR && c <= 1) ? p1 : sm_tap<CLAMP>(hrow, 1, wrap, c - 2, hw);
				e = even_inverse<WL>(lrow[c], l2, l1, hrow[c], p1);
			}
			else if (WL == AKOD_CDF53)
				e = even_inverse<WL>(lrow[c], 0, sm_tap<CLAMP>(hrow, 1, wrap, c - 1, hw), hrow[c], 0);
			else
				e = lrow[c];
			R[y * tw + 2 * c] = (int16_t)e;
		}
	}
	__syncthreads();
	// ---- H pass, odd samples (a last odd column dropped by the plus-one rule is not written)
	for (SmWalk<NT> it(hwt); it.row < th; it.next())
	{
		const int y = it.row, c = it.col;
		const int16_t* hrow = T + y * bw + hw;
		int16_t* orow = R + y * tw;
		if (2 * c + 1 < tw)
		{
			// even samples of this row sit at orow[2m]; the last one may stand in for a dropped column
			int o;
			if (WL == AKOD_DD137)
			{
				const int l1 = sm_tap<CLAMP>(orow, 2, wrap, c - 1, hw), p1 = sm_tap<CLAMP>(orow, 2, wrap, c + 1, hw);
				const int p2 = (!CLAMP && wrap == AKOD_WRAP_MIRROR && c >= hw - 2) ? l1 : sm_tap<CLAMP>(orow, 2, wrap, c + 2, hw);
				o = odd_inverse<WL>(hrow[c], orow[2 * c], l1, p1, p2);
			}
			else if (WL == AKOD_CDF53)
				o = odd_inverse<WL>(hrow[c], orow[2 * c], 0, sm_tap<CLAMP>(orow, 2, wrap, c + 1, hw), 0);
			else
				o = odd_inverse<WL>(hrow[c], orow[2 * c], 0, 0, 0);
			orow[2 * c + 1] = (int16_t)o;
		}
	}
	__syncthreads();
}

template <int NT>
__global__ void __launch_bounds__(NT, (NT <= 256) ? 4 : 1) k_unlift_small(const SmallParams p)
{
	extern __shared__ __align__(16) int16_t sm_buf[];
	int16_t* R = sm_buf;
	int16_t* T = sm_buf + p.cap;
	__shared__ uint32_t off_c[SM_MAX_LEVELS], lw[SM_MAX_LEVELS + 1], lh[SM_MAX_LEVELS + 1];

	const uint32_t img = blockIdx.x / p.channels, chn = blockIdx.x - img * p.channels;
	if (threadIdx.x == 0)
		sm_offsets(p, chn, off_c, lw, lh);
	__syncthreads();
	const int16_t* stream = p.stream + p.stream_is * img;
	const uint32_t lpn = lw[p.levels] * lh[p.levels];
	for (uint32_t i = threadIdx.x; i < lpn; i += NT)
		R[i] = __ldg(stream + chn * lpn + i);
	__syncthreads();

	for (uint32_t s = p.levels; s-- > 0;)
	{
		const int tw = (int)lw[s], th = (int)lh[s], hw = (int)lw[s + 1], hh = (int)lh[s + 1];
		const int wl = sm_level_wavelet(p.wavelet, hw, hh);
		const int16_t* in_c = stream + off_c[s];
		const int q = (int)__ldg(in_c - 1); // the decoder learns q from the lift head (misc.c:262-268)
		const bool clamp = p.wrap == AKOD_WRAP_CLAMP;
		if (wl == AKOD_DD137 && clamp)
			sm_inverse_level<AKOD_DD137, true, NT>(R, T, hw, hh, tw, th, p.wrap, q, in_c);
		else if (wl == AKOD_DD137)
			sm_inverse_level<AKOD_DD137, false, NT>(R, T, hw, hh, tw, th, p.wrap, q, in_c);
		else if (wl == AKOD_CDF53 && clamp)
			sm_inverse_level<AKOD_CDF53, true, NT>(R, T, hw, hh, tw, th, p.wrap, q, in_c);
		else if (wl == AKOD_CDF53)
			sm_inverse_level<AKOD_CDF53, false, NT>(R, T, hw, hh, tw, th, p.wrap, q, in_c);
		else
			sm_inverse_level<AKOD_HAAR, true, NT>(R, T, hw, hh, tw, th, p.wrap, q, in_c);
	}

	int16_t* out = p.planes + p.planes_is * img + p.planes_ps * chn;
	for (SmWalk<NT> it((int)p.cw0); it.row < (int)p.ch0; it.next())
		out[(uint64_t)it.row * p.planes_rs + it.col] = R[it.row * (int)p.cw0 + it.col];
}

// host side: elements each of the two shared buffers must hold for the pyramid from a cw x ch plane downwards.
// forward: 'cur' walks forward through buffer A (each level's V-highpass scratch th*2tw, then the next LL); the
// H-pass buffer holds ch rows of 2tw. inverse: R holds the 4 subbands (4*hw*hh), T 2hh rows of 2hw.
static inline uint64_t small_capacity(uint32_t cw, uint32_t ch, uint32_t levels)
{
	uint64_t at = 0, cap = 0;
	for (uint32_t s = 0; s < levels; s++)
	{
		const uint32_t tw = sm_half(cw), th = sm_half(ch);
		const uint64_t a = (uint64_t)ch * 2 * tw, b = at + (uint64_t)th * 3 * tw, c = (uint64_t)4 * tw * th;
		cap = a > cap ? a : cap;
		cap = b > cap ? b : cap;
		cap = c > cap ? c : cap;
		at += (uint64_t)th * 2 * tw;
		cw = tw;
		ch = th;
	}
	return (cap + 7) & ~(uint64_t)7;
}

// can the small kernels take the pyramid from a cw x ch plane downwards?
static inline bool small_eligible(uint32_t cw, uint32_t ch, uint32_t levels)
{
	if (levels == 0 || levels > SM_MAX_LEVELS || (uint64_t)cw * ch > SM_MAX_SAMPLES)
		return false;
	return small_capacity(cw, ch, levels) <= SM_CAP;
}
