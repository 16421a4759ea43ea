// lift.cuh -- fused 2-D integer lifting kernels (one launch per pyramid level).
//
// Forward kernel replaces, per level and channel: sLift2d (lifting.c:43-76) = ako<W>LiftH on every row
// (+ the duplicated last row when the height is odd) then ako<W>LiftV, followed by the three s2dMemcpy
// gate+quantise copies into the coefficient stream (lifting.c:154-168, :251-263).
// Inverse kernel replaces s2dUnliftHp (lifting.c:104-148): sInverseQuantization x3, ako<W>InPlaceishUnliftV x2,
// ako<W>UnliftH x2.
//
// Formulation (see oracle/ako_oracle.c, pinned against the reference): with t = ceil(n/2),
//     e(c) = x[2c],  o(c) = x[2c+1]  (o(t-1) = x[2t-2] when n is odd: "fake last")
//     CDF53 : H(c) = o(c) - (e(c) + E(c+1))/2            L(c) = e(c) + (H(c-1) + H(c))/4
//     DD137 : H(c) = o(c) + (E(c-1) + E(c+2) - 9(e(c) + E(c+1)))/16
//             L(c) = e(c) + (-H(c-2) - H(c+1) + 9(H(c-1) + H(c)))/32
//     Haar  : H = o - e, L = e
// where out-of-range taps follow the wrap mode as an index map in coefficient space
// (CLAMP/MIRROR: clamp, REPEAT: modulo t, ZERO: value 0), plus for DD137+MIRROR the substitutions
// "E(c+2) := E(c-1) for c >= t-2" and "H(c-2) := H(c+1) for c <= 1" (wavelet-dd137.c:123, :164).
// All divisions truncate toward zero; every H and L is narrowed to int16 where the reference stores it.
//
// A CTA owns a TW x TH tile of each of the four subbands. The tile of input samples plus its halo is
// staged in shared memory with the wrap mode ALREADY APPLIED by the loader (each shared-memory slot is a
// virtual coefficient index; the loader fetches the mapped real sample), so the arithmetic phases read
// plain neighbouring slots. Order is H then V on the way in, V then H on the way out.
#pragma once

#include "common.cuh"
#include "format.cuh"

#define AKOD_WRAP_CLAMP 0
#define AKOD_WRAP_MIRROR 1
#define AKOD_WRAP_REPEAT 2
#define AKOD_WRAP_ZERO 3

constexpr int LIFT_TW = 64; // coefficients per tile row (per subband)
constexpr int LIFT_TH = 32; // coefficient rows per tile
constexpr int LIFT_THREADS = 256;

template <int WL, int TW_ = LIFT_TW, int TH_ = LIFT_TH>
struct LiftGeom
{
	static constexpr int HALO = (WL == AKOD_DD137) ? 3 : (WL == AKOD_CDF53) ? 1 : 0; // coefficient slots each side
	static constexpr int HOFF = (WL == AKOD_DD137) ? 2 : (WL == AKOD_CDF53) ? 1 : 0; // H values needed left of tile
	static constexpr int HEXT = (WL == AKOD_DD137) ? 1 : 0;                           // ... and right of it
	static constexpr int EOFF = (WL == AKOD_DD137) ? 1 : 0;                           // inverse: evens left of tile
	static constexpr int EEXT = (WL == AKOD_DD137) ? 2 : (WL == AKOD_CDF53) ? 1 : 0;  // ... and right of it
	static constexpr int NS = TW_ + 2 * HALO;
	static constexpr int MS = TH_ + 2 * HALO;
	static constexpr int HBW = TW_ + HOFF + HEXT;
	static constexpr int HVH = TH_ + HOFF + HEXT;
	static constexpr int EW = TW_ + EOFF + EEXT;
	static constexpr int EH = TH_ + EOFF + EEXT;
};

// Where a CTA's tile starts. Plain launch (frame == 0): tile (blockIdx.x, blockIdx.y) of the grid of TW x TH tiles.
// The wrap modes differ only in what a tap beyond the plane's edge reads, i.e. in coefficients within three positions
// of the edge (DD 13/7: H(c) reads E(c-1..c+2), L(c) reads H(c-2..c+1); the inverse likewise). A level with another
// wrap mode than CLAMP is therefore done by the CLAMP strip kernel and then has its FRAME computed again by these
// kernels with the real wrap mode, as two launches of thin tiles anchored to the edges:
//   FRAME_ROWS  tiles of LIFT_TW x FRAME_BAND, grid (tiles across, 2): the top and the bottom FRAME_BAND rows
//   FRAME_COLS  tiles of FRAME_BAND x LIFT_TH, grid (2, tiles down):   the left and the right FRAME_BAND columns
// (w, h = the level's size in coefficients; the corners are computed twice, to the same values).
constexpr int FRAME_ROWS = 1, FRAME_COLS = 2;
constexpr int FRAME_BAND = 8;

template <int TW, int TH>
__device__ __forceinline__ void lift_tile_origin(int frame, int w, int h, int& c0, int& r0)
{
	c0 = (int)blockIdx.x * TW;
	r0 = (int)blockIdx.y * TH;
	if (frame == FRAME_ROWS && blockIdx.y != 0)
		r0 = max(h - TH, 0);
	if (frame == FRAME_COLS && blockIdx.x != 0)
		c0 = max(w - TW, 0);
}

// wrap mode as an index map; -1 means "the tap reads zero"
__device__ __forceinline__ int wrap_map(int v, int t, int wrap)
{
	if (v >= 0 && v < t)
		return v;
	if (wrap == AKOD_WRAP_CLAMP || wrap == AKOD_WRAP_MIRROR)
		return v < 0 ? 0 : t - 1;
	if (wrap == AKOD_WRAP_REPEAT)
	{
		int m = v % t;
		return m < 0 ? m + t : m;
	}
	return -1;
}

// position at which a (possibly virtual) coefficient v is evaluated
__device__ __forceinline__ int wrap_eval_pos(int v, int t, int wrap)
{
	if (wrap == AKOD_WRAP_CLAMP || wrap == AKOD_WRAP_MIRROR)
		return min(max(v, 0), t - 1);
	return v;
}

template <int WL>
__device__ __forceinline__ int16_t hp_forward(int o, int e, int l1, int p1, int p2)
{
	if (WL == AKOD_HAAR)
		return (int16_t)(o - e);
	if (WL == AKOD_CDF53)
		return (int16_t)(o - (e + p1) / 2);
	return (int16_t)(o + ((l1 + p2 - 9 * (e + p1)) / 16));
}

template <int WL>
__device__ __forceinline__ int16_t lp_forward(int e, int l2, int l1, int h, int p1)
{
	if (WL == AKOD_HAAR)
		return (int16_t)e;
	if (WL == AKOD_CDF53)
		return (int16_t)(e + (l1 + h) / 4);
	return (int16_t)(e + ((-l2 - p1 + 9 * (l1 + h)) / 32));
}

template <int WL>
__device__ __forceinline__ int16_t even_inverse(int lp, int l2, int l1, int h, int p1)
{
	if (WL == AKOD_HAAR)
		return (int16_t)lp;
	if (WL == AKOD_CDF53)
		return (int16_t)(lp - (l1 + h) / 4);
	return (int16_t)(lp - ((-l2 - p1 + 9 * (l1 + h)) / 32));
}

template <int WL>
__device__ __forceinline__ int16_t odd_inverse(int hp, int e, int l1, int p1, int p2)
{
	if (WL == AKOD_HAAR)
		return (int16_t)(e + hp);
	if (WL == AKOD_CDF53)
		return (int16_t)(hp + (e + p1) / 2);
	return (int16_t)(hp - ((l1 + p2 - 9 * (e + p1)) / 16));
}

struct LiftParams
{
	const int16_t* in; // current LL, cw x ch
	uint32_t in_rs;
	uint64_t in_ps, in_is; // plane / image stride
	int16_t* ll;           // next LL, tw x th
	uint32_t ll_rs;
	uint64_t ll_ps, ll_is;
	int16_t* stream;
	uint64_t stream_is;
	uint32_t cw, ch, tw, th;
	int wrap;
	int frame; // k_lift_level only: 0, FRAME_ROWS or FRAME_COLS (see lift_tile_origin)
	// k_lift_level only: != nullptr => level 0 of a 4-channel image read from the interleaved RGBA8 image itself (the
	// frame of a level whose inside the fused strip kernel did, lift_strip4.cuh); 'in' is unused then
	const uint8_t* rgba;
	uint64_t rgba_is, rgba_rs; // bytes between images / rows
	int rgba_color, rgba_discard;
	uint32_t channels;
	uint64_t off_c[AKOD_MAX_CHANNELS];
	int16_t q[AKOD_MAX_CHANNELS];
	int16_t g[AKOD_MAX_CHANNELS];
	uint32_t qmagic[AKOD_MAX_CHANNELS]; // ceil(2^32 / q) for q > 1
	uint32_t qmul[AKOD_MAX_CHANNELS];   // strip kernel: floor(2^qshift / q) + 1, see strip_quant
	int32_t qshift[AKOD_MAX_CHANNELS];  // 15 + ceil(log2 q)
};

// lifting.c:163 -- (v < -g || v > g) ? v / q : 0, with the truncating division done as an exact
// multiply-high (|v| <= 32768, q <= 32765  =>  |v| * (magic*q - 2^32) < 2^32, so the result is exact)
__device__ __forceinline__ int16_t gate_quantize(int v, int q, int g, uint32_t magic)
{
	if (!(v < -g || v > g))
		return 0;
	if (q <= 1)
		return (int16_t)v;
	const uint32_t a = (uint32_t)abs(v);
	const int d = (int)__umulhi(a, magic);
	return (int16_t)(v < 0 ? -d : d);
}

template <int WL, int TW = LIFT_TW, int TH = LIFT_TH>
__global__ void __launch_bounds__(LIFT_THREADS) k_lift_level(const LiftParams p)
{
	using G = LiftGeom<WL, TW, TH>;
	constexpr int XW = 2 * G::NS, XH = 2 * G::MS;

	extern __shared__ int16_t smem[];
	int16_t* X = smem;                  // XH x XW input samples (later reused for HV)
	int16_t* HB = X + XH * XW;          // XH x HBW  horizontal highpass
	int16_t* LB = HB + XH * G::HBW;     // XH x TW   horizontal lowpass
	int16_t* HV = X;                    // HVH x 2TW vertical highpass of [LB | HB]

	const uint32_t img = blockIdx.z / p.channels, chn = blockIdx.z - img * p.channels;
	const int tw = (int)p.tw, th = (int)p.th, wrap = p.wrap;
	int c0, r0;
	lift_tile_origin<TW, TH>(p.frame, tw, th, c0, r0);
	const int16_t* in = p.in + p.in_is * img + p.in_ps * chn;

	// ---- stage the tile: slot (sy,sx) <-> virtual coefficient (vr,vc) + parity; wrap applied here
	for (int i = threadIdx.x; i < XH * XW; i += LIFT_THREADS)
	{
		const int sy = i / XW, sx = i - sy * XW;
		const int mr = wrap_map(r0 - G::HALO + (sy >> 1), th, wrap);
		const int mc = wrap_map(c0 - G::HALO + (sx >> 1), tw, wrap);
		int16_t v = 0;
		if (mr >= 0 && mc >= 0)
		{
			const uint32_t y = min((uint32_t)(2 * mr + (sy & 1)), p.ch - 1); // odd height: duplicate last row
			const uint32_t x = min((uint32_t)(2 * mc + (sx & 1)), p.cw - 1); // odd width: duplicate last column
			if (p.rgba != nullptr)
			{
				// format.c:33-51, :64-135 for this one sample (k_format_fwd_rgba8x8 does the same for whole rows)
				const uint32_t px = __ldg(reinterpret_cast<const uint32_t*>(p.rgba + p.rgba_is * img + (uint64_t)y * p.rgba_rs) + x);
				int r = px & 255, g = (px >> 8) & 255, bl = (px >> 16) & 255;
				const int al = px >> 24;
				if (p.rgba_discard && al == 0)
					r = g = bl = 0;
				int16_t c0v, c1v, c2v;
				color_forward(p.rgba_color, r, g, bl, c0v, c1v, c2v);
				v = (chn == 0) ? c0v : (chn == 1) ? c1v : (chn == 2) ? c2v : (int16_t)al;
			}
			else
				v = __ldg(in + (uint64_t)y * p.in_rs + x);
		}
		X[i] = v;
	}
	__syncthreads();

	// ---- horizontal highpass for every staged row
	for (int i = threadIdx.x; i < XH * G::HBW; i += LIFT_THREADS)
	{
		const int sy = i / G::HBW, j = i - sy * G::HBW;
		const int v = c0 - G::HOFF + j;
		int16_t h = 0;
		if (!(wrap == AKOD_WRAP_ZERO && (v < 0 || v >= tw)))
		{
			const int cc = wrap_eval_pos(v, tw, wrap);
			const int16_t* row = X + sy * XW + 2 * (cc - (c0 - G::HALO));
			if (WL == AKOD_DD137)
			{
				const int l1 = row[-2], p1 = row[2];
				const int p2 = (wrap == AKOD_WRAP_MIRROR && cc >= tw - 2) ? l1 : row[4];
				h = hp_forward<WL>(row[1], row[0], l1, p1, p2);
			}
			else if (WL == AKOD_CDF53)
				h = hp_forward<WL>(row[1], row[0], 0, row[2], 0);
			else
				h = hp_forward<WL>(row[1], row[0], 0, 0, 0);
		}
		HB[i] = h;
	}
	__syncthreads();

	// ---- horizontal lowpass
	for (int i = threadIdx.x; i < XH * TW; i += LIFT_THREADS)
	{
		const int sy = i / TW, j = i - sy * TW;
		const int c = c0 + j;
		int16_t l = 0;
		if (c < tw)
		{
			const int e = X[sy * XW + 2 * (j + G::HALO)];
			const int16_t* hrow = HB + sy * G::HBW + j + G::HOFF;
			if (WL == AKOD_DD137)
			{
				const int l1 = hrow[-1], p1 = hrow[1];
				const int l2 = (wrap == AKOD_WRAP_MIRROR && c <= 1) ? p1 : hrow[-2];
				l = lp_forward<WL>(e, l2, l1, hrow[0], p1);
			}
			else if (WL == AKOD_CDF53)
				l = lp_forward<WL>(e, 0, hrow[-1], hrow[0], 0);
			else
				l = (int16_t)e;
		}
		LB[i] = l;
	}
	__syncthreads();

	// column 'col' of the horizontally transformed tile: [0,TW) lowpass half, [TW,2TW) highpass half
	auto colv = [&](int sy, int col) -> int {
		return (col < TW) ? LB[sy * TW + col] : HB[sy * G::HBW + (col - TW) + G::HOFF];
	};

	// ---- vertical highpass (overwrites X, which is dead now)
	for (int i = threadIdx.x; i < G::HVH * 2 * TW; i += LIFT_THREADS)
	{
		const int k = i / (2 * TW), col = i - k * (2 * TW);
		const int v = r0 - G::HOFF + k;
		int16_t h = 0;
		if (!(wrap == AKOD_WRAP_ZERO && (v < 0 || v >= th)))
		{
			const int rr = wrap_eval_pos(v, th, wrap);
			const int sy = 2 * (rr - (r0 - G::HALO));
			if (WL == AKOD_DD137)
			{
				const int l1 = colv(sy - 2, col), p1 = colv(sy + 2, col);
				const int p2 = (wrap == AKOD_WRAP_MIRROR && rr >= th - 2) ? l1 : colv(sy + 4, col);
				h = hp_forward<WL>(colv(sy + 1, col), colv(sy, col), l1, p1, p2);
			}
			else if (WL == AKOD_CDF53)
				h = hp_forward<WL>(colv(sy + 1, col), colv(sy, col), 0, colv(sy + 2, col), 0);
			else
				h = hp_forward<WL>(colv(sy + 1, col), colv(sy, col), 0, 0, 0);
		}
		// X and LB/HB do not overlap, so HV can be written while other threads still read LB/HB
		HV[i] = h;
	}
	__syncthreads();

	// ---- vertical lowpass + gate/quantise + stores
	const uint64_t band = (uint64_t)p.tw * p.th;
	int16_t* ll = p.ll + p.ll_is * img + p.ll_ps * chn;
	int16_t* out_c = p.stream + p.stream_is * img + p.off_c[chn];
	int16_t* out_b = out_c + band;
	int16_t* out_d = out_b + band;
	const int q = p.q[chn], g = p.g[chn];
	const uint32_t magic = p.qmagic[chn];

	// akoLiftHead{q} sits right before the C subband (lifting.c:266-267)
	if (c0 == 0 && r0 == 0 && threadIdx.x == 0)
		out_c[-1] = (int16_t)q;

	for (int i = threadIdx.x; i < TH * 2 * TW; i += LIFT_THREADS)
	{
		const int k = i / (2 * TW), col = i - k * (2 * TW);
		const int r = r0 + k;
		const int c = c0 + (col < TW ? col : col - TW);
		if (r >= th || c >= tw)
			continue;
		const int16_t* hcol = HV + (k + G::HOFF) * 2 * TW + col;
		const int e = colv(2 * (k + G::HALO), col);
		int16_t l;
		if (WL == AKOD_DD137)
		{
			const int l1 = hcol[-2 * TW], p1 = hcol[2 * TW];
			const int l2 = (wrap == AKOD_WRAP_MIRROR && r <= 1) ? p1 : hcol[-4 * TW];
			l = lp_forward<WL>(e, l2, l1, hcol[0], p1);
		}
		else if (WL == AKOD_CDF53)
			l = lp_forward<WL>(e, 0, hcol[-2 * TW], hcol[0], 0);
		else
			l = (int16_t)e;

		const uint64_t o = (uint64_t)r * p.tw + c;
		if (col < TW)
		{
			ll[(uint64_t)r * p.ll_rs + c] = l;                 // LL: next level's input
			out_c[o] = gate_quantize(hcol[0], q, g, magic);    // C: H-low / V-high
		}
		else
		{
			out_b[o] = gate_quantize(l, q, g, magic);          // B: H-high / V-low
			out_d[o] = gate_quantize(hcol[0], q, g, magic);    // D: high / high
		}
	}
}

template <int WL, int TW = LIFT_TW, int TH = LIFT_TH>
constexpr size_t lift_smem_bytes()
{
	using G = LiftGeom<WL, TW, TH>;
	return sizeof(int16_t) * ((size_t)(2 * G::MS) * (2 * G::NS) + (size_t)(2 * G::MS) * G::HBW + (size_t)(2 * G::MS) * TW);
}

// ------------------------------------------------------------------------------------------------
// inverse

struct UnliftParams
{
	const int16_t* ll; // lowpass input, hw x hh
	uint32_t ll_rs;
	uint64_t ll_ps, ll_is;
	const int16_t* stream;
	uint64_t stream_is;
	int16_t* out; // tw x th
	uint32_t out_rs;
	uint64_t out_ps, out_is;
	uint32_t hw, hh, tw, th;
	int wrap;
	int frame; // k_unlift_level only: 0, FRAME_ROWS or FRAME_COLS (see lift_tile_origin)
	uint32_t channels;
	uint64_t off_c[AKOD_MAX_CHANNELS];
};

template <int WL, int TW = LIFT_TW, int TH = LIFT_TH>
__global__ void __launch_bounds__(LIFT_THREADS) k_unlift_level(const UnliftParams p)
{
	using G = LiftGeom<WL, TW, TH>;
	constexpr int NS = G::NS, MS = G::MS;

	extern __shared__ int16_t smem[];
	int16_t* A0 = smem;             // LL   MS x NS
	int16_t* A1 = A0 + MS * NS;     // C
	int16_t* A2 = A1 + MS * NS;     // B
	int16_t* A3 = A2 + MS * NS;     // D
	int16_t* EV = A3 + MS * NS;     // 2 x EH x NS : vertically reconstructed even rows (left | right)
	int16_t* OV = EV + 2 * G::EH * NS; // 2 x TH x NS : odd rows
	int16_t* EHb = A0;              // 2TH x EW : horizontally reconstructed even samples (reuses A*)

	const uint32_t img = blockIdx.z / p.channels, chn = blockIdx.z - img * p.channels;
	const int hw = (int)p.hw, hh = (int)p.hh, wrap = p.wrap;
	int c0, r0;
	lift_tile_origin<TW, TH>(p.frame, hw, hh, c0, r0);
	const uint64_t band = (uint64_t)p.hw * p.hh;
	const int16_t* in_ll = p.ll + p.ll_is * img + p.ll_ps * chn;
	const int16_t* in_c = p.stream + p.stream_is * img + p.off_c[chn];
	// the decoder knows q only from the lift head stored right before C (misc.c:262-268, lifting.c:114-116)
	const int q = __ldg(in_c - 1);

	// ---- stage the four subband tiles (+halo), inverse quantisation fused (lifting.c:30-40)
	for (int i = threadIdx.x; i < MS * NS; i += LIFT_THREADS)
	{
		const int sr = i / NS, sc = i - sr * NS;
		const int mr = wrap_map(r0 - G::HALO + sr, hh, wrap);
		const int mc = wrap_map(c0 - G::HALO + sc, hw, wrap);
		int16_t a = 0, b = 0, c = 0, d = 0;
		if (mr >= 0 && mc >= 0)
		{
			const uint64_t o = (uint64_t)mr * p.hw + mc;
			a = __ldg(in_ll + (uint64_t)mr * p.ll_rs + mc);
			c = __ldg(in_c + o);
			b = __ldg(in_c + band + o);
			d = __ldg(in_c + 2 * band + o);
			if (q > 1)
			{
				c = (int16_t)(c * q);
				b = (int16_t)(b * q);
				d = (int16_t)(d * q);
			}
		}
		A0[i] = a;
		A1[i] = c;
		A2[i] = b;
		A3[i] = d;
	}
	__syncthreads();

	// ---- vertical: even rows, for both column halves (left = LL/C, right = B/D)
	for (int i = threadIdx.x; i < 2 * G::EH * NS; i += LIFT_THREADS)
	{
		const int side = i / (G::EH * NS);
		const int rem = i - side * (G::EH * NS);
		const int k = rem / NS, sc = rem - k * NS;
		const int v = r0 - G::EOFF + k;
		int16_t e = 0;
		if (!(wrap == AKOD_WRAP_ZERO && (v < 0 || v >= hh)))
		{
			const int rr = wrap_eval_pos(v, hh, wrap);
			const int sr = rr - (r0 - G::HALO);
			const int16_t* lo = (side ? A2 : A0) + sr * NS + sc;
			const int16_t* hi = (side ? A3 : A1) + sr * NS + sc;
			if (WL == AKOD_DD137)
			{
				const int l1 = hi[-NS], p1 = hi[NS];
				const int l2 = (wrap == AKOD_WRAP_MIRROR && rr <= 1) ? p1 : hi[-2 * NS];
				e = even_inverse<WL>(lo[0], l2, l1, hi[0], p1);
			}
			else if (WL == AKOD_CDF53)
				e = even_inverse<WL>(lo[0], 0, hi[-NS], hi[0], 0);
			else
				e = lo[0];
		}
		EV[i] = e;
	}
	__syncthreads();

	// ---- vertical: odd rows
	for (int i = threadIdx.x; i < 2 * TH * NS; i += LIFT_THREADS)
	{
		const int side = i / (TH * NS);
		const int rem = i - side * (TH * NS);
		const int k = rem / NS, sc = rem - k * NS;
		const int r = r0 + k;
		int16_t o = 0;
		if (r < hh)
		{
			const int hp = ((side ? A3 : A1) + (k + G::HALO) * NS)[sc];
			const int16_t* ev = EV + side * (G::EH * NS) + (k + G::EOFF) * NS + sc;
			if (WL == AKOD_DD137)
			{
				const int l1 = ev[-NS], p1 = ev[NS];
				const int p2 = (wrap == AKOD_WRAP_MIRROR && r >= hh - 2) ? l1 : ev[2 * NS];
				o = odd_inverse<WL>(hp, ev[0], l1, p1, p2);
			}
			else if (WL == AKOD_CDF53)
				o = odd_inverse<WL>(hp, ev[0], 0, ev[NS], 0);
			else
				o = odd_inverse<WL>(hp, ev[0], 0, 0, 0);
		}
		OV[i] = o;
	}
	__syncthreads();

	// row 'y' (0..2TH) of the vertically reconstructed tile, as (lowpass, highpass) over column slots
	auto rowl = [&](int y, int sc) -> int {
		return (y & 1) ? OV[(y >> 1) * NS + sc] : EV[((y >> 1) + G::EOFF) * NS + sc];
	};
	auto rowh = [&](int y, int sc) -> int {
		return (y & 1) ? OV[TH * NS + (y >> 1) * NS + sc] : EV[G::EH * NS + ((y >> 1) + G::EOFF) * NS + sc];
	};

	// ---- horizontal: even samples (A* are dead: EHb overlays them)
	for (int i = threadIdx.x; i < 2 * TH * G::EW; i += LIFT_THREADS)
	{
		const int y = i / G::EW, j = i - y * G::EW;
		const int v = c0 - G::EOFF + j;
		int16_t e = 0;
		if (!(wrap == AKOD_WRAP_ZERO && (v < 0 || v >= hw)))
		{
			const int cc = wrap_eval_pos(v, hw, wrap);
			const int sc = cc - (c0 - G::HALO);
			if (WL == AKOD_DD137)
			{
				const int l1 = rowh(y, sc - 1), p1 = rowh(y, sc + 1);
				const int l2 = (wrap == AKOD_WRAP_MIRROR && cc <= 1) ? p1 : rowh(y, sc - 2);
				e = even_inverse<WL>(rowl(y, sc), l2, l1, rowh(y, sc), p1);
			}
			else if (WL == AKOD_CDF53)
				e = even_inverse<WL>(rowl(y, sc), 0, rowh(y, sc - 1), rowh(y, sc), 0);
			else
				e = (int16_t)rowl(y, sc);
		}
		EHb[i] = e;
	}
	__syncthreads();

	// ---- horizontal: odd samples + stores. The sample dropped by the plus-one rule
	//      (ignore_last, lifting.c:111-112, :132) is simply not written.
	int16_t* out = p.out + p.out_is * img + p.out_ps * chn;
	for (int i = threadIdx.x; i < 2 * TH * TW; i += LIFT_THREADS)
	{
		const int y = i / TW, j = i - y * TW;
		const int c = c0 + j;
		const uint32_t oy = (uint32_t)(2 * r0 + y);
		if (c >= hw || oy >= p.th)
			continue;
		const int16_t* ev = EHb + y * G::EW + j + G::EOFF;
		const int hp = rowh(y, j + G::HALO);
		int16_t o;
		if (WL == AKOD_DD137)
		{
			const int l1 = ev[-1], p1 = ev[1];
			const int p2 = (wrap == AKOD_WRAP_MIRROR && c >= hw - 2) ? l1 : ev[2];
			o = odd_inverse<WL>(hp, ev[0], l1, p1, p2);
		}
		else if (WL == AKOD_CDF53)
			o = odd_inverse<WL>(hp, ev[0], 0, ev[1], 0);
		else
			o = odd_inverse<WL>(hp, ev[0], 0, 0, 0);

		int16_t* dst = out + (uint64_t)oy * p.out_rs + 2 * c;
		dst[0] = ev[0];
		if ((uint32_t)(2 * c + 1) < p.tw)
			dst[1] = o;
	}
}

template <int WL, int TW = LIFT_TW, int TH = LIFT_TH>
constexpr size_t unlift_smem_bytes()
{
	using G = LiftGeom<WL, TW, TH>;
	return sizeof(int16_t) * ((size_t)4 * G::MS * G::NS + (size_t)2 * G::EH * G::NS + (size_t)2 * TH * G::NS);
}
