/*
 * ako_oracle.c -- TEST INFRASTRUCTURE ONLY (see ako_oracle.h).
 *
 * Plain-C restatement of the Ako hot path. It is written from the algorithm
 * (formulas + boundary index maps), not from the reference's loop structure,
 * so that it doubles as the specification the CUDA kernels are written to.
 * Every function cites the reference lines it restates (paths under
 * /root/reference/library/). Arithmetic notes that matter for bit-exactness:
 *   - every intermediate is narrowed to int16 by wrap, exactly where the
 *     reference stores into an int16_t;
 *   - all divisions are C divisions (truncate toward zero), never shifts;
 *   - the quantiser schedule is float32 libm, same expression order.
 */
#include "ako_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum
{
	W_DD137 = 0,
	W_CDF53 = 1,
	W_HAAR = 2,
	W_NONE = 3
};
enum
{
	WRAP_CLAMP = 0,
	WRAP_MIRROR = 1,
	WRAP_REPEAT = 2,
	WRAP_ZERO = 3
};
enum
{
	COL_YCOCG = 0,
	COL_SUBTRACT_G = 1,
	COL_NONE = 2,
	COL_YCOCG_Q = 3
};
enum
{
	ST_OK = 0,
	ST_ERROR = 1,
	ST_INVALID_CHANNELS_NO = 2,
	ST_INVALID_DIMENSIONS = 3,
	ST_INVALID_TILES_DIMENSIONS = 4,
	ST_INVALID_WRAP_MODE = 5,
	ST_INVALID_WAVELET = 6,
	ST_INVALID_COLOR = 7,
	ST_INVALID_COMPRESSION = 8,
	ST_INVALID_INPUT = 9,
	ST_INVALID_CALLBACKS = 10,
	ST_INVALID_MAGIC = 11,
	ST_UNSUPPORTED_VERSION = 12,
	ST_NO_ENOUGH_MEMORY = 13,
	ST_INVALID_FLAGS = 14,
	ST_BROKEN_INPUT = 15
};

typedef int16_t i16;

/* ------------------------------------------------------------------ */
/* Synthetic image, SURVEY.md Appendix C                               */
/* ------------------------------------------------------------------ */

static uint32_t s_mix(uint32_t v)
{
	v ^= v >> 16;
	v *= 0x7feb352dU;
	v ^= v >> 15;
	v *= 0x846ca68bU;
	v ^= v >> 16;
	return v;
}

static int s_tri(uint32_t t, uint32_t period)
{
	const uint32_t p = t % period, h = period / 2;
	return (int)(p < h ? p : period - p);
}

static uint8_t s_cl(int v)
{
	return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
}

void orc_synth_rgba8(uint32_t w, uint32_t h, uint32_t seed, uint8_t* out)
{
	for (uint32_t Y = 0; Y < h; Y++)
		for (uint32_t X = 0; X < w; X++)
		{
			const uint32_t n = s_mix(X * 0x9E3779B1u ^ s_mix(Y + seed * 0x85EBCA6Bu));
			const int n0 = (int)(n & 7) - 4, n1 = (int)((n >> 8) & 7) - 4, n2 = (int)((n >> 16) & 7) - 4;
			const int blk = (int)(s_mix(((X >> 6) * 73856093u) ^ ((Y >> 6) * 19349663u) ^ seed) & 63);
			const int r = s_tri(X + 3 * seed, 509) * 255 / 254;
			const int g = s_tri(Y + 5 * seed, 383) * 255 / 191;
			const int b = s_tri(X + Y, 251) * 255 / 125;
			uint8_t* px = out + ((size_t)Y * w + X) * 4;
			px[0] = s_cl(r / 2 + blk + 32 + n0);
			px[1] = s_cl(g / 2 + blk + 32 + n1);
			px[2] = s_cl(b / 2 + (63 - blk) + 32 + n2);
			px[3] = (((X >> 7) + (Y >> 7) + seed) % 5 == 0) ? s_cl(s_tri(X, 128) * 4) : 255;
		}
}

/* ------------------------------------------------------------------ */
/* Geometry, misc.c:98-203                                             */
/* ------------------------------------------------------------------ */

/* misc.c:98-101 "divide plus one rule" = ceil(v/2) */
size_t orc_half(size_t v)
{
	return (v + 1) / 2;
}

/* number of lift levels: misc.c:138 loop condition */
size_t orc_levels(size_t w, size_t h)
{
	size_t n = 0;
	while (w > 2 && h > 2)
	{
		w = orc_half(w);
		h = orc_half(h);
		n++;
	}
	return n;
}

/* misc.c:117-149, bytes per channel */
size_t orc_tile_data_size(size_t w, size_t h)
{
	size_t bytes = 0;
	while (w > 2 && h > 2)
	{
		w = orc_half(w);
		h = orc_half(h);
		bytes += w * h * 2 * 3 + 2;
	}
	return bytes + w * h * 2;
}

/* misc.c:152-161 */
size_t orc_tile_dimension(size_t pos, size_t image_d, size_t td)
{
	if (td == 0)
		return image_d;
	if (pos + td > image_d)
		return image_d % td;
	return td;
}

/* misc.c:192-203 */
size_t orc_tiles_no(size_t w, size_t h, size_t td)
{
	if (td == 0)
		return 1;
	return ((w + td - 1) / td) * ((h + td - 1) / td);
}

/* ------------------------------------------------------------------ */
/* Quantiser schedule, quantization.c:43-98 (float32, order preserved) */
/* ------------------------------------------------------------------ */

static float s_schedule(float factor, float tile_w, float tile_h, float cur_w, float cur_h)
{
	const float area0 = sqrtf(tile_w * tile_h);
	const float area = sqrtf(cur_w * cur_h);
	const float total_lifts = log2f(area0) - 1.0F;
	const float current_lift = log2f(area) - 1.0F;
	const float linear = (current_lift / total_lifts);
	const float degrade_highs = powf(linear + 1.0F, 6.0F) / powf(2.0F, 6.0F);
	const float lg = powf(2.0F, (current_lift - 1.0F)) * degrade_highs;
	return roundf(lg * (factor / (512.0F * 0.73F)));
}

int16_t orc_quantization(int factor, int mul, size_t tile_w, size_t tile_h, size_t cur_w, size_t cur_h)
{
	if (factor <= 0)
		return 1;
	float q = s_schedule((float)factor * (float)mul, (float)tile_w, (float)tile_h, (float)cur_w, (float)cur_h);
	if (q < 1.0F)
		q = 1.0F;
	if (q > 32765.0F)
		q = 32765.0F;
	return (int16_t)q;
}

int16_t orc_gate(int factor, int mul, size_t tile_w, size_t tile_h, size_t cur_w, size_t cur_h)
{
	if (factor <= 0)
		return 0;
	float g = s_schedule((float)factor * (float)mul, (float)tile_w, (float)tile_h, (float)cur_w, (float)cur_h);
	if (g < 0.0F)
		g = 0.0F;
	if (g > 32765.0F)
		g = 32765.0F;
	return (int16_t)g;
}

/* ------------------------------------------------------------------ */
/* Format / colour, format.c                                           */
/* ------------------------------------------------------------------ */

/* format.c:30-135 */
void orc_format_forward(int discard, int color, size_t channels, size_t w, size_t h, size_t in_stride_px,
                        const uint8_t* in, int16_t* planes)
{
	const size_t plane = w * h;
	/* discard_non_visible only honoured for 2 and 4 channels (format.c:74-83) */
	const int use_discard = (discard != 0) && (channels == 2 || channels == 4);

	for (size_t y = 0; y < h; y++)
		for (size_t x = 0; x < w; x++)
		{
			const uint8_t* px = in + (y * in_stride_px + x) * channels;
			const int visible = !use_discard || px[channels - 1] != 0;
			for (size_t c = 0; c < channels; c++)
			{
				i16 v = px[c];
				if (!visible && c != channels - 1)
					v = 0;
				planes[plane * c + y * w + x] = v;
			}
		}

	if (channels < 3)
		return;

	for (size_t i = 0; i < plane; i++)
	{
		const i16 r = planes[i], g = planes[plane + i], b = planes[plane * 2 + i];
		if (color == COL_YCOCG || color == COL_YCOCG_Q)
		{
			/* format.c:94-104, :107-119 */
			const i16 co = (i16)(r - b);
			const i16 t = (i16)(b + ((r - b) / 2));
			const i16 cg = (i16)(g - t);
			i16 yy;
			if (color == COL_YCOCG)
				yy = (i16)(t + ((g - t) / 2));
			else
				yy = (i16)((t + ((g - t) / 2)) * 2);
			planes[i] = yy;
			planes[plane + i] = co;
			planes[plane * 2 + i] = cg;
		}
		else if (color == COL_SUBTRACT_G)
		{
			/* format.c:123-132 */
			planes[i] = g;
			planes[plane + i] = (i16)(r - g);
			planes[plane * 2 + i] = (i16)(b - g);
		}
	}
}

static uint8_t s_sat(i16 v)
{
	return (uint8_t)((v > 0) ? ((v < 255) ? v : 255) : 0);
}

/* format.c:138-311 */
void orc_format_inverse(int color, size_t channels, size_t w, size_t h, size_t out_stride_px, int16_t* planes,
                        uint8_t* out)
{
	const size_t plane = w * h;
	for (size_t y = 0; y < h; y++)
		for (size_t x = 0; x < w; x++)
		{
			const size_t i = y * w + x;
			uint8_t* px = out + (y * out_stride_px + x) * channels;
			size_t first_plain = 0;
			if (channels >= 3 && color != COL_NONE)
			{
				const i16 p0 = planes[i], u = planes[plane + i], v = planes[plane * 2 + i];
				i16 r, g, b;
				if (color == COL_SUBTRACT_G)
				{
					/* format.c:197-204 */
					r = (i16)(u + p0);
					g = p0;
					b = (i16)(v + p0);
				}
				else
				{
					/* format.c:142-149, :170 */
					const i16 yy = (color == COL_YCOCG_Q) ? (i16)(p0 / 2) : p0;
					const i16 t = (i16)(yy - (v / 2));
					g = (i16)(v + t);
					b = (i16)(t - (u / 2));
					r = (i16)(b + u);
				}
				px[0] = s_sat(r);
				px[1] = s_sat(g);
				px[2] = s_sat(b);
				first_plain = 3;
			}
			for (size_t c = first_plain; c < channels; c++)
				px[c] = s_sat(planes[plane * c + i]);
		}
}

/* ------------------------------------------------------------------ */
/* 1-D lifting, generic statement                                      */
/*                                                                     */
/* n samples x[0..n); t = ceil(n/2) coefficients per band.             */
/*   e(c) = x[2c]                                                      */
/*   o(c) = x[2c+1], except o(t-1) = x[2t-2] when n is odd             */
/*          ("fake_last", wavelet-cdf53.c:86-90, lifting.c:46-47)       */
/* Out-of-range taps go through s_map(): the reference's four wrap     */
/* modes are index maps in COEFFICIENT space (period t), with one      */
/* tap substitution for DD137+MIRROR (see below).                       */
/* ------------------------------------------------------------------ */

/* returns mapped index in [0,t) or -1 for "value is zero" */
static long s_map(int wrap, long v, long t)
{
	if (v >= 0 && v < t)
		return v;
	switch (wrap)
	{
	case WRAP_CLAMP:
	case WRAP_MIRROR: return (v < 0) ? 0 : t - 1;
	case WRAP_REPEAT: return ((v % t) + t) % t;
	default: return -1;
	}
}

static int s_tap(const i16* a, size_t stride, int wrap, long v, long t)
{
	const long m = s_map(wrap, v, t);
	return (m < 0) ? 0 : a[(size_t)m * stride];
}

/* Which 1-D wavelet a 2-D level really uses: lifting.c:49, :58, :67 */
static int s_level_wavelet(int wavelet, size_t tw, size_t th)
{
	if (wavelet == W_HAAR)
		return W_HAAR;
	if (wavelet == W_CDF53 || tw < 8 || th < 8)
		return W_CDF53;
	return W_DD137;
}

void orc_lift_1d(int wavelet, int wrap, size_t n, const int16_t* x, size_t xs, int16_t* lp, int16_t* hp,
                 size_t os)
{
	const long t = (long)orc_half(n);
	const int fake = (int)((size_t)t * 2 - n);

	/* gather e and o */
	i16* e = malloc(sizeof(i16) * (size_t)t * 2);
	i16* o = e + t;
	for (long c = 0; c < t; c++)
	{
		e[c] = x[(size_t)(2 * c) * xs];
		o[c] = (c == t - 1 && fake) ? e[c] : x[(size_t)(2 * c + 1) * xs];
	}

	if (wavelet == W_HAAR)
	{
		/* wavelet-haar.c:30-71 */
		for (long c = 0; c < t; c++)
		{
			lp[(size_t)c * os] = e[c];
			hp[(size_t)c * os] = (i16)(o[c] - e[c]);
		}
	}
	else if (wavelet == W_CDF53)
	{
		/* wavelet-cdf53.c:36-44, boundary taps :78-84 (even_p1) and :101-108 (hp_l1) */
		for (long c = 0; c < t; c++)
			hp[(size_t)c * os] = (i16)(o[c] - (e[c] + s_tap(e, 1, wrap, c + 1, t)) / 2);
		for (long c = 0; c < t; c++)
			lp[(size_t)c * os] = (i16)(e[c] + (s_tap(hp, os, wrap, c - 1, t) + hp[(size_t)c * os]) / 4);
	}
	else
	{
		/* wavelet-dd137.c:36-44. Boundary taps :69-77, :99-126 (HP) and :143-167, :192-203 (LP).
		 * MIRROR is not a pure index map: for c >= t-2 the +2 tap is replaced by the -1 tap
		 * (:123), and for c <= 1 the -2 tap of LP is replaced by the +1 tap (:164). */
		for (long c = 0; c < t; c++)
		{
			const int l1 = s_tap(e, 1, wrap, c - 1, t);
			const int p1 = s_tap(e, 1, wrap, c + 1, t);
			const int p2 = (wrap == WRAP_MIRROR && c >= t - 2) ? l1 : s_tap(e, 1, wrap, c + 2, t);
			hp[(size_t)c * os] = (i16)(o[c] + ((l1 + p2 - 9 * (e[c] + p1)) / 16));
		}
		for (long c = 0; c < t; c++)
		{
			const int l1 = s_tap(hp, os, wrap, c - 1, t);
			const int p1 = s_tap(hp, os, wrap, c + 1, t);
			const int l2 = (wrap == WRAP_MIRROR && c <= 1) ? p1 : s_tap(hp, os, wrap, c - 2, t);
			lp[(size_t)c * os] = (i16)(e[c] + ((-l2 - p1 + 9 * (l1 + hp[(size_t)c * os])) / 32));
		}
	}
	free(e);
}

void orc_unlift_1d(int wavelet, int wrap, size_t n, const int16_t* lp, const int16_t* hp, size_t is, int16_t* x,
                   size_t xs)
{
	const long t = (long)orc_half(n);
	const int ignore = (int)((size_t)t * 2 - n);
	i16* e = malloc(sizeof(i16) * (size_t)t * 2);
	i16* o = e + t;

	if (wavelet == W_HAAR)
	{
		/* wavelet-haar.c:74-113 */
		for (long c = 0; c < t; c++)
		{
			e[c] = lp[(size_t)c * is];
			o[c] = (i16)(lp[(size_t)c * is] + hp[(size_t)c * is]);
		}
	}
	else if (wavelet == W_CDF53)
	{
		/* wavelet-cdf53.c:46-54, :200-362 */
		for (long c = 0; c < t; c++)
			e[c] = (i16)(lp[(size_t)c * is] - (s_tap(hp, is, wrap, c - 1, t) + hp[(size_t)c * is]) / 4);
		for (long c = 0; c < t; c++)
			o[c] = (i16)(hp[(size_t)c * is] + (e[c] + s_tap(e, 1, wrap, c + 1, t)) / 2);
	}
	else
	{
		/* wavelet-dd137.c:46-54, :378-702 */
		for (long c = 0; c < t; c++)
		{
			const int l1 = s_tap(hp, is, wrap, c - 1, t);
			const int p1 = s_tap(hp, is, wrap, c + 1, t);
			const int l2 = (wrap == WRAP_MIRROR && c <= 1) ? p1 : s_tap(hp, is, wrap, c - 2, t);
			e[c] = (i16)(lp[(size_t)c * is] - ((-l2 - p1 + 9 * (l1 + hp[(size_t)c * is])) / 32));
		}
		for (long c = 0; c < t; c++)
		{
			const int l1 = s_tap(e, 1, wrap, c - 1, t);
			const int p1 = s_tap(e, 1, wrap, c + 1, t);
			const int p2 = (wrap == WRAP_MIRROR && c >= t - 2) ? l1 : s_tap(e, 1, wrap, c + 2, t);
			o[c] = (i16)(hp[(size_t)c * is] - ((l1 + p2 - 9 * (e[c] + p1)) / 16));
		}
	}

	for (long c = 0; c < t; c++)
	{
		x[(size_t)(2 * c) * xs] = e[c];
		if (!(c == t - 1 && ignore))
			x[(size_t)(2 * c + 1) * xs] = o[c];
	}
	free(e);
}

/* ------------------------------------------------------------------ */
/* Multi-level 2-D, lifting.c                                          */
/* ------------------------------------------------------------------ */

/* lifting.c:154-168 */
static i16 s_quantize(i16 v, i16 q, i16 g)
{
	if (q < 1)
		q = 1;
	return (i16)((v < -g || v > +g) ? (v / q) : 0);
}

/* Stream layout (int16 units), lifting.c:171-292 writes it back to front, misc.c:229-285 reads it:
 *   [LP ch0]..[LP chC-1]  then for level = coarsest..finest, for ch = 0..C-1: [q][C][B][D]      */
void orc_lift(const orc_settings* s, size_t channels, size_t w, size_t h, int16_t* planes, int16_t* stream)
{
	const size_t plane = w * h;
	const size_t levels = orc_levels(w, h);
	i16* tmp = malloc(sizeof(i16) * (w + 1) * (h + 1) * 2);

	/* end of stream, then walk backwards exactly like the encoder (finest level is last) */
	size_t cursor = orc_tile_data_size(w, h) * channels / 2;
	size_t cw = w, ch_ = h;

	for (size_t l = 0; l < levels; l++)
	{
		const size_t tw = orc_half(cw), th = orc_half(ch_);
		const int wl = s_level_wavelet(s->wavelet, tw, th);

		for (size_t c = channels; c-- > 0;)
		{
			const int mul = (c == 0) ? 1 : s->chroma_loss + 1; /* lifting.c:202-211 */
			const i16 q = orc_quantization(s->quantization, mul, w, h, cw, ch_);
			const i16 g = orc_gate(s->gate, mul, w, h, cw, ch_);
			i16* ll = planes + plane * c; /* dense cw x ch_ */

			/* horizontal pass (lifting.c:60-63): rows -> [L | H], row stride 2*tw */
			i16* hl = tmp;                /* H-pass lowpass  : ch_ rows x tw */
			i16* hh = tmp + ch_ * tw;     /* H-pass highpass : ch_ rows x tw */
			for (size_t r = 0; r < ch_; r++)
				orc_lift_1d(wl, s->wrap, cw, ll + r * cw, 1, hl + r * tw, hh + r * tw, 1);

			/* vertical pass (lifting.c:65): columns of both halves; the odd-height "extra row" of
			 * lifting.c:61-63 is the same fake-last rule, applied along y */
			i16* out_c = stream + (cursor - 3 * tw * th);
			i16* out_b = out_c + tw * th;
			i16* out_d = out_b + tw * th;
			i16* vl = tmp + 2 * ch_ * tw;  /* th x tw scratch */
			i16* vh = vl + th * tw;
			/* left half: LL (kept) and C */
			for (size_t x = 0; x < tw; x++)
				orc_lift_1d(wl, s->wrap, ch_, hl + x, tw, vl + x, vh + x, tw);
			for (size_t i = 0; i < tw * th; i++)
			{
				ll[i] = vl[i]; /* next level input, dense tw x th */
				out_c[i] = s_quantize(vh[i], q, g);
			}
			/* right half: B (V-low) and D (V-high) */
			for (size_t x = 0; x < tw; x++)
				orc_lift_1d(wl, s->wrap, ch_, hh + x, tw, vl + x, vh + x, tw);
			for (size_t i = 0; i < tw * th; i++)
			{
				out_b[i] = s_quantize(vl[i], q, g);
				out_d[i] = s_quantize(vh[i], q, g);
			}

			cursor -= 3 * tw * th;
			cursor -= 1;
			stream[cursor] = q; /* akoLiftHead, lifting.c:266-267 */
		}
		cw = tw;
		ch_ = th;
	}

	/* lowpasses, lifting.c:280-291 */
	for (size_t c = channels; c-- > 0;)
	{
		cursor -= cw * ch_;
		memcpy(stream + cursor, planes + plane * c, cw * ch_ * sizeof(i16));
	}
	free(tmp);
}

/* lifting.c:86-148, :295-304 with misc.c:206-288 */
void orc_unlift(const orc_settings* s, size_t channels, size_t w, size_t h, int16_t* stream, int16_t* planes)
{
	const size_t plane = w * h;
	const size_t levels = orc_levels(w, h);
	size_t dw[40], dh[40];
	dw[0] = w;
	dh[0] = h;
	for (size_t l = 0; l < levels; l++)
	{
		dw[l + 1] = orc_half(dw[l]);
		dh[l + 1] = orc_half(dh[l]);
	}

	i16* tmp = malloc(sizeof(i16) * (w + 1) * (h + 1) * 2);
	size_t cursor = 0;

	for (size_t c = 0; c < channels; c++)
	{
		memcpy(planes + plane * c, stream + cursor, dw[levels] * dh[levels] * sizeof(i16));
		cursor += dw[levels] * dh[levels];
	}

	for (size_t l = levels; l-- > 0;)
	{
		const size_t hw = dw[l + 1], hh_ = dh[l + 1]; /* subband dims */
		const size_t tw = dw[l], th = dh[l];          /* output dims  */
		const int wl = s_level_wavelet(s->wavelet, hw, hh_);

		for (size_t c = 0; c < channels; c++)
		{
			const i16 q = stream[cursor++];
			i16* hp_c = stream + cursor;
			i16* hp_b = hp_c + hw * hh_;
			i16* hp_d = hp_b + hw * hh_;
			cursor += 3 * hw * hh_;
			i16* ll = planes + plane * c;

			/* lifting.c:30-40 */
			if (q > 1)
				for (size_t i = 0; i < 3 * hw * hh_; i++)
					hp_c[i] = (i16)(hp_c[i] * q);

			/* vertical first (lifting.c:118-129), column by column; rows 0..2*hh_ */
			i16* left = tmp;                 /* 2*hh_ rows x hw */
			i16* right = tmp + 2 * hh_ * hw; /* 2*hh_ rows x hw */
			for (size_t x = 0; x < hw; x++)
			{
				orc_unlift_1d(wl, s->wrap, 2 * hh_, ll + x, hp_c + x, hw, left + x, hw);
				orc_unlift_1d(wl, s->wrap, 2 * hh_, hp_b + x, hp_d + x, hw, right + x, hw);
			}
			/* horizontal (lifting.c:131-133), only the th real rows, tw real columns */
			for (size_t r = 0; r < th; r++)
				orc_unlift_1d(wl, s->wrap, tw, left + r * hw, right + r * hw, 1, ll + r * tw, 1);
		}
	}
	free(tmp);
}

/* ------------------------------------------------------------------ */
/* Kagari, kagari.c                                                    */
/* ------------------------------------------------------------------ */

typedef struct
{
	uint8_t* out;
	uint64_t bits;
	uint64_t cap_bits;
	int count_only;
} bitw;

/* append 'len' bits of 'code' MSB first (kagari.c:59-116 produce exactly this bit order) */
static void s_put(bitw* w, uint32_t code, int len)
{
	if (!w->count_only)
		for (int i = len - 1; i >= 0; i--)
		{
			const uint64_t p = w->bits + (uint64_t)(len - 1 - i);
			if (p < w->cap_bits && ((code >> i) & 1))
				w->out[p >> 3] |= (uint8_t)(0x80u >> (p & 7));
		}
	w->bits += (uint64_t)len;
}

/* Elias gamma of a uint16 (kagari.c:38-45, :59-87); v == 0 degenerates to one 0 bit */
static void s_gamma(bitw* w, uint16_t v)
{
	int b = 0;
	for (uint16_t t = v; t > 1; t >>= 1)
		b++;
	s_put(w, v, 2 * b + 1);
}

/* kagari.c:169-173 and :214-217 (the +1, narrowed to uint16 by the callee's parameter) */
static void s_value(bitw* w, i16 x)
{
	const uint16_t zz = (uint16_t)(((int)x << 1) ^ ((int)x >> 15));
	s_gamma(w, (uint16_t)(zz + 1));
}

/* Element-rule statement of akoKagariEncode (kagari.c:228-298); see SURVEY.md 7.3 */
static void s_kagari_emit(bitw* w, size_t n, const i16* a)
{
	size_t k = 0; /* position inside the current run of equal values */
	for (size_t i = 0; i < n; i++)
	{
		k = (i > 0 && a[i] == a[i - 1]) ? k + 1 : 0;
		uint32_t c = (k == 0) ? 0 : (uint32_t)((k - 1) % 65534) + 1;
		if (c <= 2)
			s_value(w, a[i]);
		else if (c == 65534)
		{
			s_gamma(w, 65533); /* kagari.c:265-271 */
			c = 0;
		}
		const int last_of_run = (i + 1 == n) || (a[i + 1] != a[i]);
		if (last_of_run && c >= 2)
			s_gamma(w, (uint16_t)(c - 1)); /* kagari.c:275-279, :290-294 */
	}
}

uint64_t orc_kagari_bits(size_t n, const int16_t* in)
{
	bitw w = {NULL, 0, 0, 1};
	s_kagari_emit(&w, n, in);
	return w.bits;
}

size_t orc_kagari_encode(size_t n, const int16_t* in, size_t out_cap, uint8_t* out)
{
	if (n == 0 || out_cap == 0)
		return 0;
	const uint64_t bits = orc_kagari_bits(n, in);
	const uint64_t bytes = (bits + 7) / 8;
	/* the reference's "fits" rule, derived from kagari.c:65-68 and :93-107 */
	if (bytes >= out_cap)
		return 0;
	memset(out, 0, bytes);
	bitw w = {out, 0, bytes * 8, 0};
	s_kagari_emit(&w, n, in);
	return (size_t)bytes;
}

/* Bit reader that mimics the reference's 64-bit accumulator so that the number of bytes
 * "consumed" (kagari.c:119-163, :365) is reproduced, not just the values. */
typedef struct
{
	const uint8_t* cur;
	const uint8_t* end;
	uint64_t acc;
	int usage;
} bitr;

static uint16_t s_get_gamma(bitr* r, int* bits)
{
	if (r->acc == 0 || r->usage < 32)
	{
		while (r->usage < 56 && r->cur < r->end)
		{
			r->usage += 8;
			r->acc |= (uint64_t)(*r->cur) << (64 - r->usage);
			r->cur++;
		}
		if (r->acc == 0)
			return 0;
	}
	const uint32_t top = (uint32_t)(r->acc >> 32);
	const int z = top ? __builtin_clz(top) : 32;
	const int total = 2 * z + 1;
	if (total > r->usage)
		return 0;
	*bits = total;
	const uint16_t v = (uint16_t)(r->acc >> (64 - total));
	r->acc <<= total;
	r->usage -= total;
	return v;
}

size_t orc_kagari_decode(size_t n, size_t in_size, const uint8_t* in, int16_t* out)
{
	bitr r = {in, in + in_size, 0, 0};
	size_t produced = 0;
	int cn = 0;
	i16 prev = 0;
	if (n == 0 || in_size == 0)
		return 0;

	while (produced < n)
	{
		int bits = 0;
		const uint16_t u = (uint16_t)(s_get_gamma(&r, &bits) - 1);
		if (bits == 0)
			return 0;
		const i16 v = (i16)((u >> 1) ^ (uint16_t)(~(u & 1) + 1)); /* kagari.c:175-178 */
		out[produced++] = v;
		if (produced > 1 && v == prev)
		{
			if (++cn == 2)
			{
				bits = 0;
				const uint16_t len = (uint16_t)(s_get_gamma(&r, &bits) - 1);
				if (bits == 0 || produced + len > n)
					return 0;
				for (uint16_t j = 0; j < len; j++)
					out[produced++] = prev;
				cn = 0;
			}
		}
		else
		{
			prev = v;
			cn = 0;
		}
	}
	return (size_t)(r.cur - in);
}

/* ------------------------------------------------------------------ */
/* Container and whole codec                                           */
/* ------------------------------------------------------------------ */

/* head.c:34-64 */
static int s_validate(size_t channels, size_t w, size_t h, uint64_t td, int wrap, int wavelet, int color,
                      int compression)
{
	if (channels > 16)
		return ST_INVALID_CHANNELS_NO;
	if (w == 0 || h == 0 || w > 4294967295u || h > 4294967295u)
		return ST_INVALID_DIMENSIONS;
	if (td != 0 && (td < 8 || td > 2147483648u))
		return ST_INVALID_TILES_DIMENSIONS;
	if (wrap < 0 || wrap > 3)
		return ST_INVALID_WRAP_MODE;
	if (wavelet < 0 || wavelet > 3)
		return ST_INVALID_WAVELET;
	if (color < 0 || color > 3)
		return ST_INVALID_COLOR;
	if (compression < 0 || compression > 2)
		return ST_INVALID_COMPRESSION;
	return ST_OK;
}

static void s_put32(uint8_t* p, uint32_t v)
{
	p[0] = (uint8_t)v;
	p[1] = (uint8_t)(v >> 8);
	p[2] = (uint8_t)(v >> 16);
	p[3] = (uint8_t)(v >> 24);
}

static uint32_t s_get32(const uint8_t* p)
{
	return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* head.c:67-109 */
int orc_head_write(size_t channels, size_t w, size_t h, const orc_settings* s, uint8_t out[16])
{
	uint64_t code = 0;
	if (s->tiles_dimension != 0)
	{
		for (uint64_t b = s->tiles_dimension; b > 1; b >>= 1)
			code++;
		if (((uint64_t)1 << code) != s->tiles_dimension)
			return ST_INVALID_TILES_DIMENSIONS;
		code -= 2;
	}
	const int v = s_validate(channels, w, h, s->tiles_dimension, s->wrap, s->wavelet, s->color, s->compression);
	if (v != ST_OK)
		return v;
	out[0] = 'A';
	out[1] = 'k';
	out[2] = 'o';
	out[3] = 2;
	s_put32(out + 4, (uint32_t)w);
	s_put32(out + 8, (uint32_t)h);
	uint32_t flags = (uint32_t)(channels - 1);
	flags |= (uint32_t)s->wrap << 4;
	flags |= (uint32_t)s->wavelet << 6;
	flags |= (uint32_t)s->color << 8;
	flags |= (uint32_t)s->compression << 10;
	flags |= (uint32_t)code << 12;
	s_put32(out + 12, flags);
	return ST_OK;
}

/* head.c:112-169 */
int orc_head_read(const uint8_t in[16], size_t* channels, size_t* w, size_t* h, orc_settings* s)
{
	if (in[0] != 'A' || in[1] != 'k' || in[2] != 'o')
		return ST_INVALID_MAGIC;
	if (in[3] != 2)
		return ST_UNSUPPORTED_VERSION;
	const uint32_t flags = s_get32(in + 12);
	if ((flags >> 15) != 0)
		return ST_INVALID_FLAGS;
	const size_t ch = (flags & 15) + 1;
	const int wrap = (int)((flags >> 4) & 3), wavelet = (int)((flags >> 6) & 3);
	const int color = (int)((flags >> 8) & 3), compression = (int)((flags >> 10) & 3);
	uint64_t td = (flags >> 12) & 31;
	if (td != 0)
		td = (uint64_t)1 << (td + 2);
	const int v = s_validate(ch, s_get32(in + 4), s_get32(in + 8), td, wrap, wavelet, color, compression);
	if (v != ST_OK)
		return v;
	if (channels)
		*channels = ch;
	if (w)
		*w = s_get32(in + 4);
	if (h)
		*h = s_get32(in + 8);
	if (s)
	{
		s->wrap = wrap;
		s->wavelet = wavelet;
		s->color = color;
		s->compression = compression;
		s->tiles_dimension = td;
	}
	return ST_OK;
}

size_t orc_encode_bound(size_t channels, size_t w, size_t h)
{
	/* every tile block is < its int16 stream + 4, and streams sum to < 2*w*h + heads */
	return 16 + (w * h * 2 + (w + h) * 64 + 4096) * channels * 2;
}

/* encode.c:38-232 + compression.c:36-55 */
size_t orc_encode(const orc_settings* s_in, size_t channels, size_t w, size_t h, const uint8_t* in, uint8_t* out,
                  int* status)
{
	orc_settings s = *s_in;
	int st;
	/* encode.c:59-64 */
	if (s.color == COL_YCOCG && (s.quantization > 0 || s.gate > 0))
		s.color = COL_YCOCG_Q;
	else if (s.color == COL_YCOCG_Q && (s.quantization <= 0 && s.gate <= 0))
		s.color = COL_YCOCG;

	if (in == NULL)
	{
		st = ST_INVALID_INPUT;
		goto fail;
	}
	if ((st = orc_head_write(channels, w, h, &s, out)) != ST_OK)
		goto fail;

	size_t size = 16;
	const size_t td = (size_t)s.tiles_dimension;
	const size_t tiles = orc_tiles_no(w, h, td);
	size_t tx = 0, ty = 0;
	const size_t max_w = orc_tile_dimension(0, w, td), max_h = orc_tile_dimension(0, h, td);
	i16* planes = malloc(sizeof(i16) * max_w * max_h * channels);
	i16* stream = malloc(orc_tile_data_size(max_w, max_h) * channels + 64);

	for (size_t t = 0; t < tiles; t++)
	{
		const size_t tw = orc_tile_dimension(tx, w, td), th = orc_tile_dimension(ty, h, td);
		orc_format_forward(s.discard_non_visible, s.color, channels, tw, th, w, in + (w * ty + tx) * channels,
		                   planes);
		size_t data_size;
		const i16* data;
		if (s.wavelet != W_NONE)
		{
			data_size = orc_tile_data_size(tw, th) * channels;
			orc_lift(&s, channels, tw, th, planes, stream);
			data = stream;
		}
		else
		{
			data_size = tw * th * channels * 2;
			data = planes;
		}

		if (s.compression != 2)
		{
			/* compression.c:40-49: capacity is the stream size, minus the 4-byte block head */
			const size_t n = orc_kagari_encode(data_size / 2, data, data_size - 4, out + size + 4);
			if (n == 0)
			{
				free(planes);
				free(stream);
				st = ST_ERROR;
				goto fail;
			}
			s_put32(out + size, (uint32_t)n);
			size += n + 4;
		}
		else
		{
			memcpy(out + size, data, data_size);
			size += data_size;
		}

		tx += td;
		if (tx >= w)
		{
			tx = 0;
			ty += td;
		}
	}
	free(planes);
	free(stream);
	if (status)
		*status = ST_OK;
	return size;
fail:
	if (status)
		*status = st;
	return 0;
}

/* decode.c:38-250 + compression.c:58-73 */
int orc_decode(size_t in_size, const uint8_t* in, uint8_t* out, orc_settings* out_s)
{
	orc_settings s;
	memset(&s, 0, sizeof(s));
	size_t channels, w, h;
	if (in == NULL)
		return ST_INVALID_INPUT;
	int st = orc_head_read(in, &channels, &w, &h, &s);
	if (st != ST_OK)
		return st;

	const size_t td = (size_t)s.tiles_dimension;
	const size_t tiles = orc_tiles_no(w, h, td);
	const size_t max_w = orc_tile_dimension(0, w, td), max_h = orc_tile_dimension(0, h, td);
	i16* planes = malloc(sizeof(i16) * max_w * max_h * channels);
	i16* stream = malloc(orc_tile_data_size(max_w, max_h) * channels + 64);
	size_t pos = 16, tx = 0, ty = 0;
	st = ST_OK;

	for (size_t t = 0; t < tiles; t++)
	{
		const size_t tw = orc_tile_dimension(tx, w, td), th = orc_tile_dimension(ty, h, td);
		const size_t data_size =
		    (s.wavelet != W_NONE) ? orc_tile_data_size(tw, th) * channels : tw * th * channels * 2;
		i16* data = (s.wavelet != W_NONE) ? stream : planes;

		if (s.compression != 2)
		{
			if (pos + 4 > in_size)
			{
				st = ST_BROKEN_INPUT;
				break;
			}
			const size_t block = s_get32(in + pos);
			if (block == 0 || pos + 4 + block > in_size)
			{
				st = ST_BROKEN_INPUT;
				break;
			}
			const size_t used = orc_kagari_decode(data_size / 2, block, in + pos + 4, data);
			if (used == 0 || used != block)
			{
				st = ST_BROKEN_INPUT;
				break;
			}
			pos += 4 + block;
		}
		else
		{
			if (pos + data_size > in_size)
			{
				st = ST_BROKEN_INPUT;
				break;
			}
			memcpy(data, in + pos, data_size);
			pos += data_size;
		}

		if (s.wavelet != W_NONE)
			orc_unlift(&s, channels, tw, th, stream, planes);
		orc_format_inverse(s.color, channels, tw, th, w, planes, out + (w * ty + tx) * channels);

		tx += td;
		if (tx >= w)
		{
			tx = 0;
			ty += td;
		}
	}
	free(planes);
	free(stream);
	if (st == ST_OK && out_s)
		*out_s = s;
	return st;
}

/* ------------------------------------------------------------------------------------------------
 * Multi-pass ratio search: restates EncodePass of tools/akoenc.cpp:111-213 on top of orc_encode.
 * 'out' receives the blob the tool would have written; *out_q the quantisation of THAT blob (the
 * tool keeps the last pass's blob whenever its size equals the chosen bound's size, :194-210),
 * *out_passes the number of akoEncodeExt calls made. A failed pass counts as size 0, like the tool.
 * The q *= 4 loop (:151-166) overflows int in the tool when no q reaches the target; here it stops
 * once q exceeds 2^28 (every q above 32765*512 quantises identically, quantization.c:86-96). */
size_t orc_encode_pass(int ratio, const orc_settings* s_in, size_t channels, size_t w, size_t h, const uint8_t* in,
                       uint8_t* out, int* out_q, size_t* out_passes, int* status)
{
	orc_settings s = *s_in;
	size_t passes = 0;
	size_t size = 0;
	int used_q = s.quantization;

	if (ratio == 0 || s.wavelet == W_NONE || s.compression == 2 /* AKO_COMPRESSION_NONE */)
	{
		size = orc_encode(&s, channels, w, h, in, out, status);
		passes = 1;
		goto done;
	}
	if (ratio == 1)
	{
		s.quantization = 0;
		s.gate = 0;
		used_q = 0;
		size = orc_encode(&s, channels, w, h, in, out, status);
		passes = 1;
		goto done;
	}

	{
		const size_t target_size = (w * h * channels) / (size_t)ratio;
		const size_t error_margin = (target_size * 4) / 100;
		int last_q;

		s.quantization = 0;
		size_t ceil_size = orc_encode(&s, channels, w, h, in, out, status);
		last_q = 0;
		passes++;

		s.quantization = 1;
		size_t floor_size = ceil_size;
		int floor_q = 0, ceil_q = 0;
		do
		{
			s.quantization *= 4;
			ceil_size = floor_size;
			ceil_q = floor_q;
			floor_size = orc_encode(&s, channels, w, h, in, out, status);
			floor_q = s.quantization;
			last_q = s.quantization;
			passes++;
		} while (floor_size > target_size && s.quantization <= (1 << 28));

		size_t last_size = floor_size;
		while ((floor_size > ceil_size ? floor_size - ceil_size : ceil_size - floor_size) > error_margin &&
		       abs(floor_q - ceil_q) > 1)
		{
			s.quantization = (ceil_q + floor_q) / 2;
			last_size = orc_encode(&s, channels, w, h, in, out, status);
			last_q = s.quantization;
			passes++;
			if (last_size > target_size)
			{
				ceil_size = last_size;
				ceil_q = s.quantization;
			}
			else
			{
				floor_size = last_size;
				floor_q = s.quantization;
			}
		}

		const size_t df = floor_size > target_size ? floor_size - target_size : target_size - floor_size;
		const size_t dc = ceil_size > target_size ? ceil_size - target_size : target_size - ceil_size;
		const size_t chosen_size = (df < dc) ? floor_size : ceil_size;
		const int chosen_q = (df < dc) ? floor_q : ceil_q;
		if (last_size == chosen_size)
		{
			size = last_size;
			used_q = last_q;
		}
		else
		{
			s.quantization = chosen_q;
			size = orc_encode(&s, channels, w, h, in, out, status);
			used_q = chosen_q;
			passes++;
		}
	}

done:
	if (out_q)
		*out_q = used_q;
	if (out_passes)
		*out_passes = passes;
	return size;
}
