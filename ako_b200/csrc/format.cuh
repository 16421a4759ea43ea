// format.cuh -- u8 interleaved <-> int16 planar with the reversible colour transforms.
// Replaces akoFormatToPlanarI16Yuv / akoFormatToInterleavedU8Rgb (reference library/format.c:64-135, :244-311).
// Pure streaming kernels (no reuse): the bound is HBM; 4-channel images take the 128-bit path.
#pragma once

#include "common.cuh"

#define AKOD_COL_YCOCG 0
#define AKOD_COL_SUBTRACT_G 1
#define AKOD_COL_NONE 2
#define AKOD_COL_YCOCG_Q 3

// All intermediates are narrowed to int16 exactly where format.c stores into an int16_t,
// and "/ 2" is C division (toward zero), not a shift.
__device__ __forceinline__ void color_forward(int color, int r, int g, int b, int16_t& p0, int16_t& p1, int16_t& p2)
{
	if (color == AKOD_COL_YCOCG || color == AKOD_COL_YCOCG_Q)
	{
		// format.c:94-104, :107-119
		const int16_t co = (int16_t)(r - b);
		const int16_t t = (int16_t)(b + ((r - b) / 2));
		const int16_t cg = (int16_t)(g - t);
		const int y = t + ((g - t) / 2);
		p0 = (color == AKOD_COL_YCOCG) ? (int16_t)y : (int16_t)(y * 2);
		p1 = co;
		p2 = cg;
	}
	else if (color == AKOD_COL_SUBTRACT_G)
	{
		// format.c:123-132
		p0 = (int16_t)g;
		p1 = (int16_t)(r - g);
		p2 = (int16_t)(b - g);
	}
	else
	{
		p0 = (int16_t)r;
		p1 = (int16_t)g;
		p2 = (int16_t)b;
	}
}

__device__ __forceinline__ int sat_u8(int16_t v)
{
	return (v > 0) ? ((v < 255) ? v : 255) : 0;
}

__device__ __forceinline__ void color_inverse(int color, int16_t p0, int16_t u, int16_t v, int& r, int& g, int& b)
{
	if (color == AKOD_COL_YCOCG || color == AKOD_COL_YCOCG_Q)
	{
		// format.c:142-149 ; :170 for the _Q halving of the first plane
		const int16_t y = (color == AKOD_COL_YCOCG_Q) ? (int16_t)(p0 / 2) : p0;
		const int16_t t = (int16_t)(y - (v / 2));
		const int16_t gg = (int16_t)(v + t);
		const int16_t bb = (int16_t)(t - (u / 2));
		const int16_t rr = (int16_t)(bb + u);
		r = sat_u8(rr);
		g = sat_u8(gg);
		b = sat_u8(bb);
	}
	else if (color == AKOD_COL_SUBTRACT_G)
	{
		// format.c:197-204
		r = sat_u8((int16_t)(u + p0));
		g = sat_u8(p0);
		b = sat_u8((int16_t)(v + p0));
	}
	else
	{
		r = sat_u8(p0);
		g = sat_u8(u);
		b = sat_u8(v);
	}
}

// Batch member -> its interleaved u8 image (see akodBatch): a plain batch, or tiles of n_real images.
struct FmtTiles
{
	uint32_t n_real, cols, first, step, x0, y0;
};

__device__ __forceinline__ uint64_t fmt_member_offset(const FmtTiles& t, uint32_t v, uint64_t img_stride, uint64_t stride_px,
                                                      uint32_t channels)
{
	if (t.n_real == 0)
		return img_stride * v;
	const uint32_t kk = v / t.n_real, i = v - kk * t.n_real;
	const uint32_t k = t.first + kk;
	const uint32_t ky = k / t.cols, kx = k - ky * t.cols;
	return img_stride * i + ((uint64_t)(t.y0 + ky * t.step) * stride_px + (t.x0 + kx * t.step)) * channels;
}

// ---- 4 channels, rows whose width is a multiple of 8 and 16-byte aligned: 8 pixels per thread,
//      two 128-bit loads in, four 128-bit stores out.
__global__ void __launch_bounds__(256)
    k_format_fwd_rgba8x8(const uint8_t* __restrict__ in, int16_t* __restrict__ planes, uint32_t w, uint32_t h,
                         uint64_t in_stride_px, int color, int discard, uint64_t in_img_stride,
                         uint64_t planes_img_stride, const FmtTiles tiles, uint32_t pitch)
{
	const uint32_t groups_per_row = w >> 3;
	const uint64_t total = (uint64_t)groups_per_row * h;
	const uint64_t plane = (uint64_t)pitch * h;
	in += fmt_member_offset(tiles, blockIdx.y, in_img_stride, in_stride_px, 4);
	planes += planes_img_stride * blockIdx.y;

	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
	{
		const uint32_t y = (uint32_t)(i / groups_per_row);
		const uint32_t x = (uint32_t)(i - (uint64_t)y * groups_per_row) << 3;
		const uint4* src = reinterpret_cast<const uint4*>(in + ((uint64_t)y * in_stride_px + x) * 4);
		const uint4 a = __ldg(src), b = __ldg(src + 1);
		const uint32_t px[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
		int16_t o0[8], o1[8], o2[8], o3[8];
#pragma unroll
		for (int k = 0; k < 8; k++)
		{
			int r = px[k] & 255, g = (px[k] >> 8) & 255, bl = (px[k] >> 16) & 255;
			const int al = px[k] >> 24;
			if (discard && al == 0) // format.c:33-51
				r = g = bl = 0;
			color_forward(color, r, g, bl, o0[k], o1[k], o2[k]);
			o3[k] = (int16_t)al;
		}
		const uint64_t o = (uint64_t)y * pitch + x;
		*reinterpret_cast<uint4*>(planes + o) = *reinterpret_cast<const uint4*>(o0);
		*reinterpret_cast<uint4*>(planes + plane + o) = *reinterpret_cast<const uint4*>(o1);
		*reinterpret_cast<uint4*>(planes + plane * 2 + o) = *reinterpret_cast<const uint4*>(o2);
		*reinterpret_cast<uint4*>(planes + plane * 3 + o) = *reinterpret_cast<const uint4*>(o3);
	}
}

// ---- 3 channels (RGB8, the common case of real images), rows whose width is a multiple of 8: 8 pixels per thread,
//      three 64-bit loads in (24 bytes), three 128-bit plane stores out. No alpha, so no discard.
__global__ void __launch_bounds__(256)
    k_format_fwd_rgb8x8(const uint8_t* __restrict__ in, int16_t* __restrict__ planes, uint32_t w, uint32_t h,
                        uint64_t in_stride_px, int color, uint64_t in_img_stride, uint64_t planes_img_stride,
                        const FmtTiles tiles, uint32_t pitch)
{
	const uint32_t groups_per_row = w >> 3;
	const uint64_t total = (uint64_t)groups_per_row * h;
	const uint64_t plane = (uint64_t)pitch * h;
	in += fmt_member_offset(tiles, blockIdx.y, in_img_stride, in_stride_px, 3);
	planes += planes_img_stride * blockIdx.y;

	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
	{
		const uint32_t y = (uint32_t)(i / groups_per_row);
		const uint32_t x = (uint32_t)(i - (uint64_t)y * groups_per_row) << 3;
		const uint2* src = reinterpret_cast<const uint2*>(in + ((uint64_t)y * in_stride_px + x) * 3);
		const uint2 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
		const uint32_t word[6] = {a.x, a.y, b.x, b.y, c.x, c.y}; // 24 bytes: r0 g0 b0 r1 g1 b1 ...
		int16_t o0[8], o1[8], o2[8];
#pragma unroll
		for (int k = 0; k < 8; k++)
		{
			auto byte = [&](int n) { return (int)((word[n >> 2] >> (8 * (n & 3))) & 255u); };
			color_forward(color, byte(3 * k), byte(3 * k + 1), byte(3 * k + 2), o0[k], o1[k], o2[k]);
		}
		const uint64_t o = (uint64_t)y * pitch + x;
		*reinterpret_cast<uint4*>(planes + o) = *reinterpret_cast<const uint4*>(o0);
		*reinterpret_cast<uint4*>(planes + plane + o) = *reinterpret_cast<const uint4*>(o1);
		*reinterpret_cast<uint4*>(planes + plane * 2 + o) = *reinterpret_cast<const uint4*>(o2);
	}
}

__global__ void __launch_bounds__(256)
    k_format_inv_rgb8x8(const int16_t* __restrict__ planes, uint8_t* __restrict__ out, uint32_t w, uint32_t h,
                        uint64_t out_stride_px, int color, uint64_t planes_img_stride, uint64_t out_img_stride,
                        const FmtTiles tiles, uint32_t pitch)
{
	const uint32_t groups_per_row = w >> 3;
	const uint64_t total = (uint64_t)groups_per_row * h;
	const uint64_t plane = (uint64_t)pitch * h;
	planes += planes_img_stride * blockIdx.y;
	out += fmt_member_offset(tiles, blockIdx.y, out_img_stride, out_stride_px, 3);

	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
	{
		const uint32_t y = (uint32_t)(i / groups_per_row);
		const uint32_t x = (uint32_t)(i - (uint64_t)y * groups_per_row) << 3;
		const uint64_t o = (uint64_t)y * pitch + x;
		int16_t p0[8], p1[8], p2[8];
		*reinterpret_cast<uint4*>(p0) = __ldg(reinterpret_cast<const uint4*>(planes + o));
		*reinterpret_cast<uint4*>(p1) = __ldg(reinterpret_cast<const uint4*>(planes + plane + o));
		*reinterpret_cast<uint4*>(p2) = __ldg(reinterpret_cast<const uint4*>(planes + plane * 2 + o));
		uint32_t word[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
		for (int k = 0; k < 8; k++)
		{
			int r, g, b;
			color_inverse(color, p0[k], p1[k], p2[k], r, g, b);
			const int comp[3] = {r, g, b};
#pragma unroll
			for (int n = 0; n < 3; n++)
			{
				const int at = 3 * k + n;
				word[at >> 2] |= (uint32_t)comp[n] << (8 * (at & 3));
			}
		}
		uint2* dst = reinterpret_cast<uint2*>(out + ((uint64_t)y * out_stride_px + x) * 3);
		dst[0] = make_uint2(word[0], word[1]);
		dst[1] = make_uint2(word[2], word[3]);
		dst[2] = make_uint2(word[4], word[5]);
	}
}

// ---- any channel count / any width: one pixel per thread
__global__ void __launch_bounds__(256)
    k_format_fwd_generic(const uint8_t* __restrict__ in, int16_t* __restrict__ planes, uint32_t channels, uint32_t w,
                         uint32_t h, uint64_t in_stride_px, int color, int discard, uint64_t in_img_stride,
                         uint64_t planes_img_stride, const FmtTiles tiles, uint32_t pitch)
{
	const uint64_t plane = (uint64_t)pitch * h, pixels = (uint64_t)w * h;
	in += fmt_member_offset(tiles, blockIdx.y, in_img_stride, in_stride_px, channels);
	planes += planes_img_stride * blockIdx.y;
	// discard_non_visible is only honoured for 2 and 4 channels (format.c:74-83)
	const bool use_discard = discard && (channels == 2 || channels == 4);

	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < pixels; j += (uint64_t)gridDim.x * blockDim.x)
	{
		const uint32_t y = (uint32_t)(j / w);
		const uint32_t x = (uint32_t)(j - (uint64_t)y * w);
		const uint64_t i = (uint64_t)y * pitch + x;
		const uint8_t* px = in + ((uint64_t)y * in_stride_px + x) * channels;
		const bool visible = !use_discard || px[channels - 1] != 0;
		if (channels >= 3)
		{
			int r = px[0], g = px[1], b = px[2];
			if (!visible)
				r = g = b = 0;
			int16_t p0, p1, p2;
			color_forward(color, r, g, b, p0, p1, p2);
			planes[i] = p0;
			planes[plane + i] = p1;
			planes[plane * 2 + i] = p2;
			for (uint32_t c = 3; c < channels; c++)
				planes[plane * c + i] = (int16_t)((visible || c == channels - 1) ? px[c] : 0);
		}
		else
		{
			for (uint32_t c = 0; c < channels; c++)
				planes[plane * c + i] = (int16_t)((visible || c == channels - 1) ? px[c] : 0);
		}
	}
}

__global__ void __launch_bounds__(256)
    k_format_inv_rgba8x8(const int16_t* __restrict__ planes, uint8_t* __restrict__ out, uint32_t w, uint32_t h,
                         uint64_t out_stride_px, int color, uint64_t planes_img_stride, uint64_t out_img_stride,
                         const FmtTiles tiles, uint32_t pitch)
{
	const uint32_t groups_per_row = w >> 3;
	const uint64_t total = (uint64_t)groups_per_row * h;
	const uint64_t plane = (uint64_t)pitch * h;
	planes += planes_img_stride * blockIdx.y;
	out += fmt_member_offset(tiles, blockIdx.y, out_img_stride, out_stride_px, 4);

	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
	{
		const uint32_t y = (uint32_t)(i / groups_per_row);
		const uint32_t x = (uint32_t)(i - (uint64_t)y * groups_per_row) << 3;
		const uint64_t o = (uint64_t)y * pitch + x;
		int16_t p0[8], p1[8], p2[8], p3[8];
		*reinterpret_cast<uint4*>(p0) = __ldg(reinterpret_cast<const uint4*>(planes + o));
		*reinterpret_cast<uint4*>(p1) = __ldg(reinterpret_cast<const uint4*>(planes + plane + o));
		*reinterpret_cast<uint4*>(p2) = __ldg(reinterpret_cast<const uint4*>(planes + plane * 2 + o));
		*reinterpret_cast<uint4*>(p3) = __ldg(reinterpret_cast<const uint4*>(planes + plane * 3 + o));
		uint32_t px[8];
#pragma unroll
		for (int k = 0; k < 8; k++)
		{
			int r, g, b;
			color_inverse(color, p0[k], p1[k], p2[k], r, g, b);
			px[k] = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)sat_u8(p3[k]) << 24);
		}
		uint4* dst = reinterpret_cast<uint4*>(out + ((uint64_t)y * out_stride_px + x) * 4);
		dst[0] = make_uint4(px[0], px[1], px[2], px[3]);
		dst[1] = make_uint4(px[4], px[5], px[6], px[7]);
	}
}

__global__ void __launch_bounds__(256)
    k_format_inv_generic(const int16_t* __restrict__ planes, uint8_t* __restrict__ out, uint32_t channels, uint32_t w,
                         uint32_t h, uint64_t out_stride_px, int color, uint64_t planes_img_stride,
                         uint64_t out_img_stride, const FmtTiles tiles, uint32_t pitch)
{
	const uint64_t plane = (uint64_t)pitch * h, pixels = (uint64_t)w * h;
	planes += planes_img_stride * blockIdx.y;
	out += fmt_member_offset(tiles, blockIdx.y, out_img_stride, out_stride_px, channels);

	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < pixels; j += (uint64_t)gridDim.x * blockDim.x)
	{
		const uint32_t y = (uint32_t)(j / w);
		const uint32_t x = (uint32_t)(j - (uint64_t)y * w);
		const uint64_t i = (uint64_t)y * pitch + x;
		uint8_t* px = out + ((uint64_t)y * out_stride_px + x) * channels;
		uint32_t first = 0;
		if (channels >= 3)
		{
			int r, g, b;
			color_inverse(color, planes[i], planes[plane + i], planes[plane * 2 + i], r, g, b);
			px[0] = (uint8_t)r;
			px[1] = (uint8_t)g;
			px[2] = (uint8_t)b;
			first = 3;
		}
		for (uint32_t c = first; c < channels; c++)
			px[c] = (uint8_t)sat_u8(planes[plane * c + i]);
	}
}
