/*
 * ako_host.c -- host side of libako_b200, in C like the reference.
 *
 * Holds everything that is not per-pixel work: defaults and status strings (reference misc.c:30-95),
 * version getters (version.c), the 16-byte container header (head.c), tile geometry (misc.c:98-203), the
 * float quantiser schedule (quantization.c:43-98, evaluated on the host only, same expression order, no
 * fast-math) and the orchestration of akoEncodeExt / akoDecodeExt (encode.c, decode.c). The per-pixel
 * work is issued to the CUDA layer through ako_device.h. There is no CPU implementation of any stage:
 * if the CUDA layer cannot run, every entry point fails with AKO_ERROR / AKO_NO_ENOUGH_MEMORY.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "ako_b200.h"
#include "ako_device.h"

#define AKO_API __attribute__((visibility("default")))

struct akoB200Context
{
	akodContext* dev;
	/* small plan cache: full tile, right edge, bottom edge, corner */
	akodPlan plan[4];
	struct akoSettings plan_settings[4];
	int plan_valid[4];
	int plan_next;
	int mail_pending; /* an upload kernel may still be reading the pinned mailbox (see upload_words) */
};

/* ------------------------------------------------------------------------------------------------ */
/* defaults, strings, versions                                                                       */

AKO_API struct akoSettings akoDefaultSettings(void)
{
	struct akoSettings s;
	memset(&s, 0, sizeof(s));
	s.wavelet = AKO_WAVELET_DD137;
	s.color = AKO_COLOR_YCOCG;
	s.wrap = AKO_WRAP_CLAMP;
	s.compression = AKO_COMPRESSION_KAGARI;
	s.tiles_dimension = 0;
	s.quantization = 16;
	s.gate = 0;
	s.chroma_loss = 1;
	s.discard_non_visible = 0;
	return s;
}

AKO_API struct akoCallbacks akoDefaultCallbacks(void)
{
	struct akoCallbacks c;
	memset(&c, 0, sizeof(c));
	c.malloc = malloc;
	c.realloc = realloc;
	c.free = free;
	return c;
}

AKO_API void akoDefaultFree(void* p)
{
	free(p);
}

AKO_API const char* akoStatusString(enum akoStatus status)
{
	static const char* const text[] = {
	    "Everything Ok!",
	    "Something went wrong",
	    "Invalid channels number",
	    "Invalid dimensions",
	    "Invalid tiles dimensions",
	    "Invalid wrap mode",
	    "Invalid wavelet transformation",
	    "Invalid color transformation",
	    "Invalid compression method",
	    "Invalid input",
	    "Invalid callbacks",
	    "Invalid magic (not an Ako file)",
	    "Unsupported version",
	    "No enough memory",
	    "Invalid flags",
	    "Broken input/premature end",
	};
	if ((int)status < 0 || (size_t)status >= sizeof(text) / sizeof(text[0]))
		return "Unknown status code";
	return text[status];
}

AKO_API int akoVersionMajor(void)
{
	return AKO_VERSION_MAJOR;
}
AKO_API int akoVersionMinor(void)
{
	return AKO_VERSION_MINOR;
}
AKO_API int akoVersionPatch(void)
{
	return AKO_VERSION_PATCH;
}
AKO_API int akoFormatVersion(void)
{
	return AKO_FORMAT_VERSION;
}

/* ------------------------------------------------------------------------------------------------ */
/* container header, head.c                                                                          */

static enum akoStatus validate(size_t channels, size_t w, size_t h, size_t tiles, int wrap, int wavelet, int color,
                               int compression)
{
	/* same order as head.c:38-63: the first failing property decides the status */
	if (channels > AKO_MAX_CHANNELS)
		return AKO_INVALID_CHANNELS_NO;
	if (w == 0 || h == 0 || w > AKO_MAX_WIDTH || h > AKO_MAX_HEIGHT)
		return AKO_INVALID_DIMENSIONS;
	if (tiles != 0 && (tiles < AKO_MIN_TILES_DIMENSION || tiles > AKO_MAX_TILES_DIMENSION))
		return AKO_INVALID_TILES_DIMENSIONS;
	if (wrap < AKO_WRAP_CLAMP || wrap > AKO_WRAP_ZERO)
		return AKO_INVALID_WRAP_MODE;
	if (wavelet < AKO_WAVELET_DD137 || wavelet > AKO_WAVELET_NONE)
		return AKO_INVALID_WAVELET_TRANSFORMATION;
	if (color < AKO_COLOR_YCOCG || color > AKO_COLOR_YCOCG_Q)
		return AKO_INVALID_COLOR_TRANSFORMATION;
	if (compression < AKO_COMPRESSION_KAGARI || compression > AKO_COMPRESSION_NONE)
		return AKO_INVALID_COMPRESSION_METHOD;
	return AKO_OK;
}

static void store_le32(uint8_t* p, uint32_t v)
{
	p[0] = (uint8_t)(v);
	p[1] = (uint8_t)(v >> 8);
	p[2] = (uint8_t)(v >> 16);
	p[3] = (uint8_t)(v >> 24);
}

static uint32_t load_le32(const uint8_t* p)
{
	return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* head.c:67-109. Written byte by byte, so it is endian-neutral. */
static enum akoStatus head_write(size_t channels, size_t w, size_t h, const struct akoSettings* s, uint8_t out[16])
{
	uint32_t tiles_code = 0;
	if (s->tiles_dimension != 0)
	{
		size_t log2 = 0;
		for (size_t b = s->tiles_dimension; b > 1; b >>= 1)
			log2++;
		if (((size_t)1 << log2) != s->tiles_dimension)
			return AKO_INVALID_TILES_DIMENSIONS;
		tiles_code = (uint32_t)(log2 - 2); /* 8 -> 1 */
	}

	const enum akoStatus st = validate(channels, w, h, s->tiles_dimension, (int)s->wrap, (int)s->wavelet,
	                                   (int)s->color, (int)s->compression);
	if (st != AKO_OK)
		return st;

	out[0] = 'A';
	out[1] = 'k';
	out[2] = 'o';
	out[3] = AKO_FORMAT_VERSION;
	store_le32(out + 4, (uint32_t)w);
	store_le32(out + 8, (uint32_t)h);
	store_le32(out + 12, (uint32_t)(channels - 1) | ((uint32_t)s->wrap << 4) | ((uint32_t)s->wavelet << 6) |
	                         ((uint32_t)s->color << 8) | ((uint32_t)s->compression << 10) | (tiles_code << 12));
	return AKO_OK;
}

static enum akoStatus check_sizes(size_t channels, size_t w, size_t h, size_t td);

/* head.c:112-169 */
static enum akoStatus head_read(const uint8_t in[16], size_t* channels, size_t* w, size_t* h, struct akoSettings* s)
{
	if (in[0] != 'A' || in[1] != 'k' || in[2] != 'o')
		return AKO_INVALID_MAGIC;
	if (in[3] != AKO_FORMAT_VERSION)
		return AKO_UNSUPPORTED_VERSION;

	const uint32_t flags = load_le32(in + 12);
	if ((flags >> 15) != 0)
		return AKO_INVALID_FLAGS;

	const size_t ch = (size_t)(flags & 15u) + 1;
	const int wrap = (int)((flags >> 4) & 3u);
	const int wavelet = (int)((flags >> 6) & 3u);
	const int color = (int)((flags >> 8) & 3u);
	const int compression = (int)((flags >> 10) & 3u);
	size_t tiles = (flags >> 12) & 31u;
	if (tiles != 0)
		tiles = (size_t)1 << (tiles + 2);

	enum akoStatus st = validate(ch, load_le32(in + 4), load_le32(in + 8), tiles, wrap, wavelet, color, compression);
	if (st == AKO_OK)
		st = check_sizes(ch, load_le32(in + 4), load_le32(in + 8), tiles);
	if (st != AKO_OK)
		return st;

	*channels = ch;
	*w = load_le32(in + 4);
	*h = load_le32(in + 8);
	s->wrap = (enum akoWrap)wrap;
	s->wavelet = (enum akoWavelet)wavelet;
	s->color = (enum akoColor)color;
	s->compression = (enum akoCompression)compression;
	s->tiles_dimension = tiles;
	return AKO_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* geometry, misc.c:98-203                                                                           */

static size_t half_up(size_t v) /* akoDividePlusOneRule, misc.c:98-101 */
{
	return (v >> 1) + (v & 1);
}

static size_t tile_data_size(size_t w, size_t h) /* akoTileDataSize, misc.c:117-149; bytes per channel */
{
	size_t bytes = 0;
	while (w > 2 && h > 2)
	{
		w = half_up(w);
		h = half_up(h);
		bytes += w * h * 3 * sizeof(int16_t) + sizeof(int16_t);
	}
	return bytes + w * h * sizeof(int16_t);
}

static size_t tile_dimension(size_t pos, size_t image_d, size_t tiles) /* akoTileDimension, misc.c:152-161 */
{
	if (tiles == 0)
		return image_d;
	if (pos + tiles > image_d)
		return image_d % tiles;
	return tiles;
}

static size_t tiles_count(size_t w, size_t h, size_t tiles) /* akoImageTilesNo, misc.c:192-203 */
{
	if (tiles == 0)
		return 1;
	return ((w / tiles) + (w % tiles != 0)) * ((h / tiles) + (h % tiles != 0));
}

/* Sizes that the work areas and the kernels' 32-bit element indices are built on, checked before anything is
 * allocated or launched: a crafted header (w = h = 2^31 ...) must fail here, not wrap a product. A tile's
 * coefficient stream has to stay below 2^32 values (Kagari run positions) and a plane below 2^31 samples. */
static enum akoStatus check_sizes(size_t channels, size_t w, size_t h, size_t td)
{
	size_t plane, image, tiles_x, tiles_y, tiles;
	if (__builtin_mul_overflow(w, h, &plane) || __builtin_mul_overflow(plane, channels, &image) ||
	    image > ((size_t)1 << 46))
		return AKO_NO_ENOUGH_MEMORY;
	const size_t tw = tile_dimension(0, w, td), th = tile_dimension(0, h, td);
	if (tw * th >= ((size_t)1 << 31) || tile_data_size(tw, th) * channels / 2 >= ((size_t)1 << 32))
		return AKO_NO_ENOUGH_MEMORY;
	tiles_x = (td == 0) ? 1 : (w / td) + (w % td != 0);
	tiles_y = (td == 0) ? 1 : (h / td) + (h % td != 0);
	if (__builtin_mul_overflow(tiles_x, tiles_y, &tiles) || tiles > ((size_t)1 << 28))
		return AKO_NO_ENOUGH_MEMORY;
	return AKO_OK;
}

AKO_API size_t akoB200StreamSize(size_t channels, size_t tile_w, size_t tile_h)
{
	return tile_data_size(tile_w, tile_h) * channels;
}

/* ------------------------------------------------------------------------------------------------ */
/* quantiser schedule, quantization.c:43-98 -- float32, host only                        */

static float schedule(float factor, float tile_w, float tile_h, float current_w, float current_h)
{
	const float side0 = sqrtf(tile_w * tile_h);
	const float side = sqrtf(current_w * current_h);
	const float lifts_total = log2f(side0) - 1.0F;
	const float lift_now = log2f(side) - 1.0F;
	const float ramp = lift_now / lifts_total;
	const float highs = powf(ramp + 1.0F, 6.0F) / powf(2.0F, 6.0F);
	const float step = powf(2.0F, lift_now - 1.0F) * highs;
	return roundf(step * (factor / (512.0F * 0.73F)));
}

static int16_t level_quantization(int factor, int mul, size_t tw, size_t th, size_t cw, size_t ch)
{
	if (factor <= 0)
		return 1;
	float q = schedule((float)factor * (float)mul, (float)tw, (float)th, (float)cw, (float)ch);
	q = (q < 1.0F) ? 1.0F : q;
	q = (q > 32765.0F) ? 32765.0F : q;
	return (int16_t)q;
}

static int16_t level_gate(int factor, int mul, size_t tw, size_t th, size_t cw, size_t ch)
{
	if (factor <= 0)
		return 0;
	float g = schedule((float)factor * (float)mul, (float)tw, (float)th, (float)cw, (float)ch);
	g = (g < 0.0F) ? 0.0F : g;
	g = (g > 32765.0F) ? 32765.0F : g;
	return (int16_t)g;
}

/* Builds the integer description of one tile: level dimensions, the wavelet each level really uses
 * (lifting.c:49-75), q and gate per level and channel (lifting.c:197-211) and the offset of every
 * subband in the coefficient stream (lifting.c:179, :251-267, :280-285 / misc.c:229-285). */
static void build_plan(akodPlan* p, const struct akoSettings* s, size_t channels, size_t w, size_t h)
{
	memset(p, 0, sizeof(*p));
	p->w = (uint32_t)w;
	p->h = (uint32_t)h;
	p->channels = (uint32_t)channels;
	p->wrap = (int32_t)s->wrap;
	p->wavelet = (s->wavelet == AKO_WAVELET_HAAR) ? AKOD_HAAR : (s->wavelet == AKO_WAVELET_CDF53) ? AKOD_CDF53 : AKOD_DD137;

	size_t cw = w, ch = h;
	uint32_t levels = 0;
	while (cw > 2 && ch > 2)
	{
		akodLevel* L = &p->level[levels];
		L->cw = (uint32_t)cw;
		L->ch = (uint32_t)ch;
		L->tw = (uint32_t)half_up(cw);
		L->th = (uint32_t)half_up(ch);
		if (s->wavelet == AKO_WAVELET_HAAR)
			L->wavelet = AKOD_HAAR;
		else if (s->wavelet == AKO_WAVELET_CDF53 || L->tw < 8 || L->th < 8)
			L->wavelet = AKOD_CDF53;
		else
			L->wavelet = AKOD_DD137;
		for (size_t c = 0; c < channels; c++)
		{
			const int mul = (c == 0) ? 1 : s->chroma_loss + 1;
			L->q[c] = level_quantization(s->quantization, mul, w, h, cw, ch);
			L->g[c] = level_gate(s->gate, mul, w, h, cw, ch);
		}
		cw = L->tw;
		ch = L->th;
		levels++;
	}
	p->levels = levels;
	p->lp_w = (uint32_t)cw;
	p->lp_h = (uint32_t)ch;

	/* stream order: lowpasses, then coarsest..finest level, channels ascending inside a level */
	uint64_t cursor = 0;
	for (size_t c = 0; c < channels; c++)
	{
		p->off_lp[c] = cursor;
		cursor += (uint64_t)cw * ch;
	}
	for (uint32_t l = levels; l-- > 0;)
	{
		akodLevel* L = &p->level[l];
		for (size_t c = 0; c < channels; c++)
		{
			cursor += 1; /* akoLiftHead */
			L->off_c[c] = cursor;
			cursor += 3 * (uint64_t)L->tw * L->th;
		}
	}
	p->stream_len = cursor;
}

static int same_settings_for_plan(const struct akoSettings* a, const struct akoSettings* b)
{
	return a->wavelet == b->wavelet && a->wrap == b->wrap && a->quantization == b->quantization &&
	       a->gate == b->gate && a->chroma_loss == b->chroma_loss;
}

static const akodPlan* get_plan(akoB200Context* ctx, const struct akoSettings* s, size_t channels, size_t w, size_t h)
{
	for (int i = 0; i < 4; i++)
		if (ctx->plan_valid[i] && ctx->plan[i].w == w && ctx->plan[i].h == h && ctx->plan[i].channels == channels &&
		    same_settings_for_plan(&ctx->plan_settings[i], s))
			return &ctx->plan[i];
	const int slot = ctx->plan_next;
	ctx->plan_next = (ctx->plan_next + 1) & 3;
	build_plan(&ctx->plan[slot], s, channels, w, h);
	ctx->plan_settings[slot] = *s;
	ctx->plan_valid[slot] = 1;
	return &ctx->plan[slot];
}

/* ------------------------------------------------------------------------------------------------ */
/* contexts                                                                                          */

static enum akoStatus from_dev(int rc)
{
	if (rc == AKOD_OK)
		return AKO_OK;
	return (rc == AKOD_NOMEM) ? AKO_NO_ENOUGH_MEMORY : AKO_ERROR;
}

static int env_device(void)
{
	const char* e = getenv("AKO_CUDA_DEVICE");
	return (e != NULL && *e != '\0') ? atoi(e) : 0;
}

/* The devices the host-pointer API spreads its work over: $AKO_CUDA_DEVICES = "0,1,..." (one process, several
 * GPUs: images and chunks of a batch are independent), else the single $AKO_CUDA_DEVICE (default 0). */
#define MAX_DEVICES 16
static int env_devices(int list[MAX_DEVICES])
{
	const char* e = getenv("AKO_CUDA_DEVICES");
	int n = 0;
	if (e != NULL)
		while (*e != '\0' && n < MAX_DEVICES)
		{
			char* end = NULL;
			const long d = strtol(e, &end, 10);
			if (end == e)
				break;
			if (d >= 0 && d < 1024)
				list[n++] = (int)d;
			e = (*end == ',') ? end + 1 : end;
			if (*end != ',' && *end != '\0')
				break;
		}
	if (n == 0)
		list[n++] = env_device();
	return n;
}

AKO_API akoB200Context* akoB200ContextCreate(int device, enum akoStatus* out_status)
{
	akoB200Context* ctx = calloc(1, sizeof(*ctx));
	enum akoStatus st = AKO_OK;
	if (ctx == NULL)
		st = AKO_NO_ENOUGH_MEMORY;
	else
	{
		st = from_dev(akod_context_create(device < 0 ? env_device() : device, &ctx->dev));
		if (st != AKO_OK)
		{
			free(ctx);
			ctx = NULL;
		}
	}
	if (out_status != NULL)
		*out_status = st;
	return ctx;
}

AKO_API void akoB200ContextDestroy(akoB200Context* ctx)
{
	if (ctx == NULL)
		return;
	akod_context_destroy(ctx->dev);
	free(ctx);
}

AKO_API void* akoB200ContextStream(akoB200Context* ctx)
{
	return akod_stream(ctx->dev);
}

AKO_API enum akoStatus akoB200Synchronize(akoB200Context* ctx)
{
	return from_dev(akod_sync(ctx->dev));
}

AKO_API void* akoB200DeviceAlloc(akoB200Context* ctx, size_t bytes)
{
	return akod_alloc(ctx->dev, bytes);
}

AKO_API void akoB200DeviceFree(akoB200Context* ctx, void* p)
{
	akod_free(ctx->dev, p);
}

AKO_API void* akoB200PinnedAlloc(size_t bytes)
{
	return akod_pinned_alloc(bytes);
}

AKO_API void akoB200PinnedFree(void* p)
{
	akod_pinned_free(p);
}

AKO_API enum akoStatus akoB200CopyToDevice(akoB200Context* ctx, void* d_dst, const void* src, size_t bytes)
{
	return from_dev(akod_h2d(ctx->dev, d_dst, src, bytes));
}

AKO_API enum akoStatus akoB200CopyToHost(akoB200Context* ctx, void* dst, const void* d_src, size_t bytes)
{
	return from_dev(akod_d2h(ctx->dev, dst, d_src, bytes));
}

/* Pinned realloc needs the old size: keep it in a 64-byte prefix (keeps 64-byte alignment). */
#define PIN_PREFIX 64

/* cudaHostAlloc / cudaFreeHost cost milliseconds for image-sized blocks, so freed blocks are kept in a small
 * cache (prefix word 0 = user size, word 1 = capacity) and handed out again to requests they fit. */
#define PIN_CACHE_SLOTS 512
#define PIN_CACHE_MAX_BYTES ((size_t)6 << 30)
static pthread_mutex_t g_pin_lock = PTHREAD_MUTEX_INITIALIZER;
static uint8_t* g_pin_cache[PIN_CACHE_SLOTS];
static size_t g_pin_cached_bytes = 0;

static void* pinned_malloc(size_t bytes)
{
	uint8_t* raw = NULL;
	pthread_mutex_lock(&g_pin_lock);
	{
		int best = -1;
		for (int i = 0; i < PIN_CACHE_SLOTS; i++)
			if (g_pin_cache[i] != NULL)
			{
				const size_t cap = ((size_t*)g_pin_cache[i])[1];
				if (cap >= bytes && cap <= bytes * 2 + 4096 &&
				    (best < 0 || cap < ((size_t*)g_pin_cache[best])[1]))
					best = i;
			}
		if (best >= 0)
		{
			raw = g_pin_cache[best];
			g_pin_cache[best] = NULL;
			g_pin_cached_bytes -= ((size_t*)raw)[1];
		}
	}
	pthread_mutex_unlock(&g_pin_lock);

	if (raw == NULL)
	{
		/* blobs of one image shape differ by a few percent: round up so that they share cached blocks */
		const size_t grain = (bytes < ((size_t)1 << 16)) ? 4096 : (size_t)1 << 16;
		const size_t cap = (bytes + grain - 1) & ~(grain - 1);
		raw = akod_pinned_alloc(cap + PIN_PREFIX);
		if (raw == NULL)
			return NULL;
		((size_t*)raw)[1] = cap;
	}
	((size_t*)raw)[0] = bytes;
	return raw + PIN_PREFIX;
}

static void pinned_free(void* p)
{
	if (p == NULL)
		return;
	uint8_t* raw = (uint8_t*)p - PIN_PREFIX;
	const size_t cap = ((size_t*)raw)[1];
	pthread_mutex_lock(&g_pin_lock);
	{
		/* a free slot, else the place of the smallest cached block when that one is smaller (page-locking cost
		 * grows with size: the big blocks are the ones worth keeping) */
		int slot = -1, smallest = -1;
		for (int i = 0; i < PIN_CACHE_SLOTS; i++)
		{
			if (g_pin_cache[i] == NULL)
			{
				slot = i;
				break;
			}
			if (smallest < 0 || ((size_t*)g_pin_cache[i])[1] < ((size_t*)g_pin_cache[smallest])[1])
				smallest = i;
		}
		if (slot < 0 && smallest >= 0 && ((size_t*)g_pin_cache[smallest])[1] < cap)
			slot = smallest;
		if (slot >= 0)
		{
			uint8_t* evicted = g_pin_cache[slot];
			const size_t evicted_cap = (evicted != NULL) ? ((size_t*)evicted)[1] : 0;
			if (g_pin_cached_bytes - evicted_cap + cap <= PIN_CACHE_MAX_BYTES)
			{
				g_pin_cache[slot] = raw;
				g_pin_cached_bytes += cap - evicted_cap;
				raw = evicted;
			}
		}
	}
	pthread_mutex_unlock(&g_pin_lock);
	if (raw != NULL)
		akod_pinned_free(raw);
}

static void* pinned_realloc(void* p, size_t bytes)
{
	if (p == NULL)
		return pinned_malloc(bytes);
	const size_t old = *(size_t*)((uint8_t*)p - PIN_PREFIX);
	void* n = pinned_malloc(bytes);
	if (n == NULL)
		return NULL;
	memcpy(n, p, old < bytes ? old : bytes);
	pinned_free(p);
	return n;
}

AKO_API struct akoCallbacks akoB200PinnedCallbacks(void)
{
	struct akoCallbacks c;
	memset(&c, 0, sizeof(c));
	c.malloc = pinned_malloc;
	c.realloc = pinned_realloc;
	c.free = pinned_free;
	return c;
}

AKO_API void akoB200ProfileEnable(akoB200Context* ctx, int enable)
{
	akod_profile_enable(ctx->dev, enable);
}

AKO_API void akoB200ProfileReset(akoB200Context* ctx)
{
	akod_profile_reset(ctx->dev);
}

AKO_API size_t akoB200ProfileGet(akoB200Context* ctx, size_t cap, const char** names, uint64_t* launches,
                                 double* total_ms)
{
	return akod_profile_get(ctx->dev, cap, names, launches, total_ms);
}

AKO_API size_t akoB200ProfileGetBytes(akoB200Context* ctx, size_t cap, uint64_t* bytes)
{
	return akod_profile_get_bytes(ctx->dev, cap, bytes);
}

AKO_API uint64_t akoB200LaunchCount(akoB200Context* ctx)
{
	return akod_launch_count(ctx->dev);
}

/* Contexts used by the host-pointer API: created lazily, one per concurrent caller, recycled. */
#define POOL_MAX 64
static pthread_mutex_t g_pool_lock = PTHREAD_MUTEX_INITIALIZER;
static akoB200Context* g_pool[POOL_MAX];
static int g_pool_len = 0;

static akoB200Context* pool_acquire_on(int device, enum akoStatus* st)
{
	akoB200Context* ctx = NULL;
	pthread_mutex_lock(&g_pool_lock);
	for (int i = 0; i < g_pool_len; i++)
		if (akod_device_index(g_pool[i]->dev) == device)
		{
			ctx = g_pool[i];
			g_pool[i] = g_pool[--g_pool_len];
			break;
		}
	pthread_mutex_unlock(&g_pool_lock);
	if (ctx == NULL)
	{
		ctx = akoB200ContextCreate(device, st);
		/* the pool serves host-pointer calls: their waits are PCIe copies, and their callers may outnumber the cores */
		if (ctx != NULL && getenv("AKO_B200_SPIN_SYNC") == NULL)
			akod_set_blocking_sync(ctx->dev, 1);
	}
	return ctx;
}

/* single-image calls: concurrent callers are dealt over the listed devices in turn */
static akoB200Context* pool_acquire(enum akoStatus* st)
{
	static unsigned next = 0;
	int list[MAX_DEVICES];
	const int n = env_devices(list);
	const unsigned turn = (n > 1) ? __atomic_fetch_add(&next, 1u, __ATOMIC_RELAXED) : 0u;
	return pool_acquire_on(list[turn % (unsigned)n], st);
}

static void pool_release(akoB200Context* ctx)
{
	pthread_mutex_lock(&g_pool_lock);
	if (g_pool_len < POOL_MAX)
	{
		g_pool[g_pool_len++] = ctx;
		ctx = NULL;
	}
	pthread_mutex_unlock(&g_pool_lock);
	if (ctx != NULL)
		akoB200ContextDestroy(ctx);
}

/* Small uint64 arrays to/from the device through the context's pinned mailbox (64 KiB): truly
 * asynchronous uploads, and read-backs that do not stage through pageable memory. */
#define MAILBOX_WORDS 8192

static enum akoStatus upload_words(akoB200Context* ctx, uint64_t* d_dst, const uint64_t* src, size_t words)
{
	enum akoStatus st = AKO_OK;
	uint64_t* mail = akod_mailbox(ctx->dev);
	for (size_t o = 0; o < words && st == AKO_OK; o += MAILBOX_WORDS)
	{
		const size_t m = (words - o < MAILBOX_WORDS) ? words - o : MAILBOX_WORDS;
		if (o != 0 || ctx->mail_pending)
			st = from_dev(akod_sync(ctx->dev)); /* the previous chunk, or an earlier upload that no read-back has followed */
		if (st == AKO_OK)
		{
			memcpy(mail, src + o, sizeof(uint64_t) * m);
			st = from_dev(akod_copy_words(ctx->dev, d_dst + o, mail, m)); /* a kernel reads the pinned mailbox */
			ctx->mail_pending = 1;
		}
	}
	return st;
}

static enum akoStatus download_words(akoB200Context* ctx, uint64_t* dst, const uint64_t* d_src, size_t words)
{
	enum akoStatus st = AKO_OK;
	uint64_t* mail = akod_mailbox(ctx->dev);
	for (size_t o = 0; o < words && st == AKO_OK; o += MAILBOX_WORDS)
	{
		const size_t m = (words - o < MAILBOX_WORDS) ? words - o : MAILBOX_WORDS;
		st = from_dev(akod_copy_words(ctx->dev, mail, d_src + o, m)); /* a kernel writes the pinned mailbox */
		if (st == AKO_OK)
			st = from_dev(akod_sync(ctx->dev));
		ctx->mail_pending = 0;
		if (st == AKO_OK)
			memcpy(dst + o, mail, sizeof(uint64_t) * m);
	}
	return st;
}

/* ------------------------------------------------------------------------------------------------ */
/* single stages                                                                                     */

static size_t align_up(size_t v, size_t a)
{
	return (v + a - 1) / a * a;
}

static enum akoStatus scratch_for(akoB200Context* ctx, size_t channels, size_t w, size_t h, size_t n, int16_t** out,
                                  size_t* stride)
{
	void* p = NULL;
	*stride = align_up(align_up(half_up(w), 8) * half_up(h) * channels, 8); /* rows padded to 8 elements (akod_lift) */
	const enum akoStatus st = from_dev(akod_workspace(ctx->dev, AKOD_WS_SCRATCH, *stride * n * sizeof(int16_t) + 64, &p));
	*out = p;
	return st;
}

AKO_API enum akoStatus akoB200FormatForward(akoB200Context* ctx, const struct akoSettings* s, size_t channels, size_t w,
                                            size_t h, size_t in_stride_px, const void* d_in, int16_t* d_planes)
{
	return from_dev(akod_format_forward(ctx->dev, s->discard_non_visible, (int)s->color, (uint32_t)channels, (uint32_t)w,
	                                    (uint32_t)h, in_stride_px, d_in, d_planes, NULL));
}

AKO_API enum akoStatus akoB200FormatInverse(akoB200Context* ctx, enum akoColor color, size_t channels, size_t w,
                                            size_t h, size_t out_stride_px, const int16_t* d_planes, void* d_out)
{
	return from_dev(akod_format_inverse(ctx->dev, (int)color, (uint32_t)channels, (uint32_t)w, (uint32_t)h,
	                                    out_stride_px, d_planes, d_out, NULL));
}

AKO_API enum akoStatus akoB200Lift(akoB200Context* ctx, const struct akoSettings* s, size_t channels, size_t w, size_t h,
                                   int16_t* d_planes, int16_t* d_stream)
{
	if (s->wavelet == AKO_WAVELET_NONE || channels == 0 || channels > AKO_MAX_CHANNELS)
		return AKO_ERROR;
	int16_t* scratch;
	size_t stride;
	const enum akoStatus st = scratch_for(ctx, channels, w, h, 1, &scratch, &stride);
	if (st != AKO_OK)
		return st;
	if (w % 8 != 0)
	{
		/* the cores feed the lifting planes whose rows are padded to 8 elements: so does the stage (same kernels) */
		const size_t pitch = align_up(w, 8);
		void* padded;
		akodBatch b;
		memset(&b, 0, sizeof(b));
		b.n = 1;
		b.planes_pitch = (uint32_t)pitch;
		enum akoStatus st2 = from_dev(akod_workspace(ctx->dev, AKOD_WS_PLANES, pitch * h * channels * 2 + 64, &padded));
		if (st2 == AKO_OK)
			st2 = from_dev(akod_copy_strided(ctx->dev, padded, pitch * 2, d_planes, w * 2, w * 2, h * channels));
		if (st2 != AKO_OK)
			return st2;
		return from_dev(akod_lift(ctx->dev, get_plan(ctx, s, channels, w, h), padded, scratch, d_stream, &b));
	}
	return from_dev(akod_lift(ctx->dev, get_plan(ctx, s, channels, w, h), d_planes, scratch, d_stream, NULL));
}

AKO_API enum akoStatus akoB200Unlift(akoB200Context* ctx, const struct akoSettings* s, size_t channels, size_t w,
                                     size_t h, const int16_t* d_stream, int16_t* d_planes)
{
	if (s->wavelet == AKO_WAVELET_NONE || channels == 0 || channels > AKO_MAX_CHANNELS)
		return AKO_ERROR;
	int16_t* scratch;
	size_t stride;
	const enum akoStatus st = scratch_for(ctx, channels, w, h, 1, &scratch, &stride);
	if (st != AKO_OK)
		return st;
	if (w % 8 != 0)
	{
		const size_t pitch = align_up(w, 8);
		void* padded;
		akodBatch b;
		memset(&b, 0, sizeof(b));
		b.n = 1;
		b.planes_pitch = (uint32_t)pitch;
		enum akoStatus st2 = from_dev(akod_workspace(ctx->dev, AKOD_WS_PLANES, pitch * h * channels * 2 + 64, &padded));
		if (st2 == AKO_OK)
			st2 = from_dev(akod_unlift(ctx->dev, get_plan(ctx, s, channels, w, h), d_stream, padded, scratch, &b));
		if (st2 == AKO_OK)
			st2 = from_dev(akod_copy_strided(ctx->dev, d_planes, w * 2, padded, pitch * 2, w * 2, h * channels));
		return st2;
	}
	return from_dev(akod_unlift(ctx->dev, get_plan(ctx, s, channels, w, h), d_stream, d_planes, scratch, NULL));
}

AKO_API size_t akoB200KagariEncode(akoB200Context* ctx, size_t n_values, const int16_t* d_in, void* d_out,
                                   size_t out_capacity, enum akoStatus* out_status)
{
	enum akoStatus st = AKO_OK;
	size_t result = 0;
	void* small = NULL;
	uint64_t* mail = akod_mailbox(ctx->dev);

	if (n_values == 0 || out_capacity < 4)
		st = AKO_ERROR;
	if (st == AKO_OK)
		st = from_dev(akod_workspace(ctx->dev, AKOD_WS_SMALL, 4096, &small));
	if (st == AKO_OK)
		st = from_dev(akod_kagari_encode(ctx->dev, n_values, d_in, 0, d_out, 0, out_capacity & ~(size_t)3, small, 1));
	if (st == AKO_OK)
		st = from_dev(akod_d2h(ctx->dev, mail, small, sizeof(uint64_t)));
	if (st == AKO_OK)
		st = from_dev(akod_sync(ctx->dev));
	if (st == AKO_OK)
	{
		const uint64_t bytes = (mail[0] + 7) / 8;
		/* the reference's "fits" rule: kagari.c:65-68, :93-107 */
		if (bytes < out_capacity && bytes <= (out_capacity & ~(size_t)3))
			result = (size_t)bytes;
	}
	if (out_status != NULL)
		*out_status = st;
	return result;
}

AKO_API size_t akoB200KagariDecode(akoB200Context* ctx, size_t n_values, size_t in_size, const void* d_in,
                                   int16_t* d_out, enum akoStatus* out_status)
{
	enum akoStatus st = AKO_OK;
	size_t result = 0;
	void* small = NULL;
	uint64_t words[2] = {0, in_size};
	uint64_t answer = 0;

	if (n_values == 0 || in_size == 0)
		st = AKO_ERROR;
	if (st == AKO_OK)
		st = from_dev(akod_workspace(ctx->dev, AKOD_WS_SMALL, 4096, &small));
	uint64_t* d_words = small;
	if (st == AKO_OK)
		st = upload_words(ctx, d_words, words, 2);
	if (st == AKO_OK)
		st = from_dev(akod_kagari_decode(ctx->dev, n_values, d_in, d_words, d_words + 1, in_size, d_out, 0, d_words + 2, 1));
	if (st == AKO_OK)
		st = download_words(ctx, &answer, d_words + 2, 1);
	if (st == AKO_OK)
		result = (size_t)answer;
	if (out_status != NULL)
		*out_status = st;
	return result;
}

/* ------------------------------------------------------------------------------------------------ */
/* whole codec                                                                                       */

static void fire(const struct akoCallbacks* c, akoB200Context* ctx, size_t tile, size_t tiles, enum akoEvent e)
{
	/* Events bracket GPU work, so the stream is drained first: the caller's stopwatch
	 * (tools/benchmark.hpp:73-90 in the reference) then measures the stage it names. */
	if (c != NULL && c->events != NULL)
	{
		akod_sync(ctx->dev);
		c->events(tile, tiles, e, c->events_data);
	}
}

AKO_API size_t akoB200EncodeBound(const struct akoSettings* s_in, size_t channels, size_t w, size_t h)
{
	const struct akoSettings s = (s_in != NULL) ? *s_in : akoDefaultSettings();
	const size_t td = s.tiles_dimension;
	size_t bound = 16;
	size_t tx = 0, ty = 0;
	const size_t tiles = tiles_count(w, h, td);
	for (size_t t = 0; t < tiles; t++)
	{
		const size_t tw = tile_dimension(tx, w, td), th = tile_dimension(ty, h, td);
		const size_t data = (s.wavelet != AKO_WAVELET_NONE) ? tile_data_size(tw, th) * channels : tw * th * channels * 2;
		bound += align_up(data, 4) + 4;
		tx += td;
		if (tx >= w)
		{
			tx = 0;
			ty += td;
		}
	}
	return bound;
}

/* ---- tiles as batch members ------------------------------------------------------------------------------
 * The reference walks the tiles of an image one after the other (encode.c:115-205, decode.c:113-230); every tile is
 * an independent block. Here the tiles of one shape are members of ONE batch: the full-size tiles (group 0), the
 * right edge column (1), the bottom edge row (2) and the corner (3) are at most four passes of the batched kernels,
 * whatever the number of tiles. Member v of a pass = kk * n + i is tile k0 + kk of the group in image i. */
#define MEMBERS_PER_PASS 32768 /* batch members ride in gridDim.y (<= 65535) */

struct tile_groups
{
	akodTiles T;
	size_t tiles;         /* per image */
	size_t tw[4], th[4];  /* tile shape of the group */
	size_t count[4];      /* tiles of the group per image */
	size_t data[4];       /* bytes of one tile's coefficient stream (all channels): what a block may not reach */
	uint32_t cols[4], x0[4], y0[4];
	size_t planes_stride[4], scratch_stride[4], stream_stride[4]; /* int16 elements per member */
	uint32_t planes_pitch[4];                                     /* row pitch of the planes, 0 = dense */
	size_t bits_slots;    /* tiles * n */
	size_t blocks_bytes;
};

static void make_groups(struct tile_groups* G, const struct akoSettings* s, size_t channels, size_t w, size_t h, size_t n)
{
	const size_t td = s->tiles_dimension;
	memset(G, 0, sizeof(*G));
	const size_t full_x = (td == 0) ? 1 : w / td, full_y = (td == 0) ? 1 : h / td;
	const size_t ex = (td != 0 && w % td != 0), ey = (td != 0 && h % td != 0);
	G->T.tiles_x = (uint32_t)(full_x + ex);
	G->T.tiles_y = (uint32_t)(full_y + ey);
	G->T.full_x = (uint32_t)full_x;
	G->T.full_y = (uint32_t)full_y;
	G->T.n_images = (uint32_t)n;
	G->tiles = (full_x + ex) * (full_y + ey);
	const size_t fw = (td == 0) ? w : td, fh = (td == 0) ? h : td;
	const size_t ew = (td == 0) ? 0 : w % td, eh = (td == 0) ? 0 : h % td;
	const size_t gw[4] = {fw, ew, fw, ew}, gh[4] = {fh, fh, eh, eh};
	const size_t gc[4] = {full_x * full_y, ex ? full_y : 0, ey ? full_x : 0, (ex && ey) ? 1 : 0};
	const size_t gcols[4] = {full_x, 1, full_x, 1};
	const size_t gx0[4] = {0, full_x * td, 0, full_x * td}, gy0[4] = {0, 0, full_y * td, full_y * td};
	uint64_t blocks = 0, bits = 0;
	for (int g = 0; g < 4; g++)
	{
		G->tw[g] = gw[g];
		G->th[g] = gh[g];
		G->count[g] = gc[g];
		G->cols[g] = (uint32_t)(gcols[g] ? gcols[g] : 1);
		G->x0[g] = (uint32_t)gx0[g];
		G->y0[g] = (uint32_t)gy0[g];
		if (gc[g] == 0)
			continue;
		const size_t data = (s->wavelet != AKO_WAVELET_NONE) ? tile_data_size(gw[g], gh[g]) * channels
		                                                     : gw[g] * gh[g] * channels * 2;
		G->data[g] = data;
		/* planes that feed the lifting have their rows padded to 8 elements (every row on a 16-byte boundary: the strip
		 * kernels fetch rows with the TMA engine); without a wavelet the planes ARE the tile's data and stay dense */
		G->planes_pitch[g] = (s->wavelet != AKO_WAVELET_NONE) ? (uint32_t)align_up(gw[g], 8) : 0;
		G->planes_stride[g] = align_up((G->planes_pitch[g] ? G->planes_pitch[g] : gw[g]) * gh[g] * channels, 8);
		G->scratch_stride[g] = align_up(align_up(half_up(gw[g]), 8) * half_up(gh[g]) * channels, 8);
		G->stream_stride[g] = align_up(data / 2, 8);
		/* what the packer may write: the reference accepts a block of up to data - 5 bytes (bytes < data - 4,
		 * compression.c:40-49); the packer stores whole words, and the tile's region (align_up(data, 16)) holds
		 * the word that byte data - 5 lies in. The exact rule is applied on the host from the bit counts. */
		G->T.group_cap[g] = (s->compression == AKO_COMPRESSION_NONE) ? data : (data >= 6) ? align_up(data - 5, 4) : 0;
		G->T.group_stride[g] = align_up(data, 16);
		G->T.group_base[g] = blocks;
		G->T.bits_base[g] = bits;
		blocks += (uint64_t)gc[g] * n * G->T.group_stride[g];
		bits += (uint64_t)gc[g] * n;
	}
	G->bits_slots = (size_t)bits;
	G->blocks_bytes = (size_t)blocks;
}

/* most members one pass of group g may take */
static size_t group_pass_tiles(const struct tile_groups* G, int g, size_t n, int tile_by_tile)
{
	if (tile_by_tile)
		return 1;
	size_t kc = MEMBERS_PER_PASS / n;
	kc = (kc < 1) ? 1 : kc;
	return (kc > G->count[g]) ? G->count[g] : kc;
}

static void group_batch(akodBatch* b, const struct tile_groups* G, int g, size_t k0, size_t kc, size_t n, size_t image_stride,
                        size_t td)
{
	memset(b, 0, sizeof(*b));
	b->n = (uint32_t)(kc * n);
	b->in_stride = image_stride;
	b->planes_stride = G->planes_stride[g];
	b->planes_pitch = G->planes_pitch[g];
	b->scratch_stride = G->scratch_stride[g];
	b->stream_stride = G->stream_stride[g];
	if (td != 0)
	{
		b->n_real = (uint32_t)n;
		b->tile_cols = G->cols[g];
		b->tile_first = (uint32_t)k0;
		b->tile_step = (uint32_t)td;
		b->tile_x0 = G->x0[g];
		b->tile_y0 = G->y0[g];
	}
}

/* Same-shape batch core. n images at d_in + i*in_stride -> n blobs at d_out + i*out_stride.
 * 'cb' only for events (may be NULL). */
/* Ratio search support (akoB200EncodeRatio): the UNQUANTISED coefficient stream of every tile of ONE image is kept
 * on the device, so that a probe with another quantisation only re-quantises it and measures the Kagari length. */
enum
{
	SC_FILL = 1, /* format + lift with q = 1, gate = 0 into the cache, nothing else */
	SC_USE = 2   /* take the tile's stream from the cache (re-quantised with the call's settings) */
};
struct search_cache
{
	int mode;
	int size_only;  /* SC_USE or no cache: stop after the Kagari length scan, report sizes, write no blob */
	int16_t* d_raw; /* device, every tile's stream back to back (8-element aligned) */
};

static void resolve_color(struct akoSettings* s) /* encode.c:59-64 */
{
	if (s->color == AKO_COLOR_YCOCG && (s->quantization > 0 || s->gate > 0))
		s->color = AKO_COLOR_YCOCG_Q;
	else if (s->color == AKO_COLOR_YCOCG_Q && (s->quantization <= 0 && s->gate <= 0))
		s->color = AKO_COLOR_YCOCG;
}

static size_t encode_core_ex(akoB200Context* ctx, const struct akoCallbacks* cb, const struct akoSettings* s_in,
                             size_t channels, size_t w, size_t h, size_t n, const uint8_t* d_in, size_t in_stride,
                             uint8_t* d_out, size_t out_stride, size_t out_capacity, size_t* out_sizes,
                             enum akoStatus* out_status, const struct search_cache* sc)
{
	enum akoStatus st = AKO_OK;
	struct akoSettings s = (s_in != NULL) ? *s_in : akoDefaultSettings();
	size_t done = 0;
	const int fill = (sc != NULL && sc->mode == SC_FILL);
	const int use = (sc != NULL && sc->mode == SC_USE);
	const int size_only = (sc != NULL && sc->size_only);

	if (!fill) /* a cache fill names the colour it wants */
		resolve_color(&s);
	if (sc != NULL && (n != 1 || s.wavelet == AKO_WAVELET_NONE || s.compression == AKO_COMPRESSION_NONE))
	{
		st = AKO_ERROR;
		goto done;
	}

	if (d_in == NULL)
	{
		st = AKO_INVALID_INPUT;
		goto done;
	}

	uint8_t head[16];
	if ((st = head_write(channels, w, h, &s, head)) != AKO_OK)
		goto done;
	if (channels == 0)
	{
		st = AKO_INVALID_CHANNELS_NO; /* the reference would write flags 0xFFFFFFFF; refuse instead */
		goto done;
	}
	if (s.wavelet == AKO_WAVELET_NONE && s.compression != AKO_COMPRESSION_NONE)
	{
		/* The reference compresses akoTileDataSize() bytes of a buffer it only formatted w*h*2 bytes of
		 * (compression.c:40 vs encode.c:127): its output depends on uninitialised memory. */
		st = AKO_ERROR;
		goto done;
	}
	if ((st = check_sizes(channels, w, h, s.tiles_dimension)) != AKO_OK)
		goto done;
	if (!fill && !size_only && out_capacity < akoB200EncodeBound(&s, channels, w, h))
	{
		st = AKO_NO_ENOUGH_MEMORY;
		goto done;
	}

	const size_t td = s.tiles_dimension;
	struct tile_groups G;
	make_groups(&G, &s, channels, w, h, n);
	const size_t tiles = G.tiles;
	/* a caller's stopwatch (events) and the ratio search's per-tile cache see the tiles one after the other, as the
	 * reference does them; everybody else gets one pass per tile shape */
	const int tile_by_tile = (cb != NULL && cb->events != NULL) || sc != NULL;
	if (n > MEMBERS_PER_PASS)
	{
		st = AKO_ERROR; /* the batch entry points split larger batches */
		goto done;
	}

	void *planes, *scratch, *stream, *blocks, *small;
	{
		size_t need_planes = 0, need_scratch = 0, need_stream = 0;
		for (int g = 0; g < 4; g++)
			if (G.count[g] != 0)
			{
				const size_t members = group_pass_tiles(&G, g, n, tile_by_tile) * n;
				need_planes = (G.planes_stride[g] * members > need_planes) ? G.planes_stride[g] * members : need_planes;
				need_scratch = (G.scratch_stride[g] * members > need_scratch) ? G.scratch_stride[g] * members : need_scratch;
				need_stream = (G.stream_stride[g] * members > need_stream) ? G.stream_stride[g] * members : need_stream;
			}
		if ((st = from_dev(akod_workspace(ctx->dev, AKOD_WS_PLANES, need_planes * 2 + 64, &planes))) != AKO_OK ||
		    (st = from_dev(akod_workspace(ctx->dev, AKOD_WS_SCRATCH, need_scratch * 2 + 64, &scratch))) != AKO_OK ||
		    (st = from_dev(akod_workspace(ctx->dev, AKOD_WS_STREAM, need_stream * 2 + 64, &stream))) != AKO_OK ||
		    (st = from_dev(akod_workspace(ctx->dev, AKOD_WS_BLOCKS, G.blocks_bytes + 64, &blocks))) != AKO_OK ||
		    (st = from_dev(akod_workspace(ctx->dev, AKOD_WS_SMALL, sizeof(uint64_t) * (G.bits_slots + n) + 64, &small))) != AKO_OK)
			goto done;
	}
	uint64_t* d_bits = small;              /* laid out as akodTiles says */
	uint64_t* d_total = d_bits + G.bits_slots; /* [n] */

	size_t cache_cursor = 0; /* elements; the search cache keeps the tiles in raster order */
	/* work list: per-tile mode walks the tiles in raster order, one tile of all n images per pass; otherwise every
	 * shape group is cut into passes of at most MEMBERS_PER_PASS members */
	for (size_t item = 0; st == AKO_OK; item++)
	{
		uint32_t g = 0, k0 = 0;
		size_t kc = 0, t_event = 0;
		if (tile_by_tile)
		{
			if (item >= tiles)
				break;
			akod_tile_locate(&G.T, (uint32_t)item, &g, &k0);
			kc = 1;
			t_event = item;
		}
		else
		{
			/* item enumerates (group, pass) pairs */
			size_t left = item;
			int found = 0;
			for (int gg = 0; gg < 4 && !found; gg++)
			{
				if (G.count[gg] == 0)
					continue;
				const size_t per = group_pass_tiles(&G, gg, n, 0);
				const size_t passes = (G.count[gg] + per - 1) / per;
				if (left < passes)
				{
					g = (uint32_t)gg;
					k0 = (uint32_t)(left * per);
					kc = (G.count[gg] - k0 < per) ? G.count[gg] - k0 : per;
					found = 1;
				}
				else
					left -= passes;
			}
			if (!found)
				break;
		}
		const size_t tw = G.tw[g], th = G.th[g];
		const size_t members = kc * n;
		akodBatch batch;
		group_batch(&batch, &G, (int)g, k0, kc, n, in_stride, td);
		int16_t* cached = (sc != NULL && sc->d_raw != NULL) ? sc->d_raw + cache_cursor : NULL;
		cache_cursor += align_up(G.data[g] / 2, 8);

		/* level 0 of the lifting may read the RGBA8 image itself (colour/format fused into the kernel) */
		int fuse = 0;
		if (!use && s.wavelet != AKO_WAVELET_NONE)
			fuse = akod_format_lift_fuses(ctx->dev, (uint32_t)channels, (uint32_t)tw, (uint32_t)th, w, d_in,
			                              get_plan(ctx, &s, channels, tw, th), planes, scratch, fill ? cached : stream, &batch);
		if (!use)
		{
			fire(cb, ctx, t_event, tiles, AKO_EVENT_FORMAT_START);
			if (!fuse)
				st = from_dev(akod_format_forward(ctx->dev, s.discard_non_visible, (int)s.color, (uint32_t)channels,
				                                  (uint32_t)tw, (uint32_t)th, w, d_in, planes, &batch));
			fire(cb, ctx, t_event, tiles, AKO_EVENT_FORMAT_END);
		}

		const int16_t* data = planes;
		uint64_t data_stride = G.planes_stride[g];
		if (st == AKO_OK && s.wavelet != AKO_WAVELET_NONE)
		{
			fire(cb, ctx, t_event, tiles, AKO_EVENT_WAVELET_START);
			if (use)
				st = from_dev(akod_requantize(ctx->dev, get_plan(ctx, &s, channels, tw, th), cached, stream));
			else if (fuse)
				st = from_dev(akod_format_lift(ctx->dev, s.discard_non_visible, (int)s.color, (uint32_t)channels, (uint32_t)tw,
				                               (uint32_t)th, w, d_in, get_plan(ctx, &s, channels, tw, th), planes, scratch,
				                               fill ? cached : stream, &batch));
			else
				st = from_dev(akod_lift(ctx->dev, get_plan(ctx, &s, channels, tw, th), planes, scratch,
				                        fill ? cached : stream, &batch));
			fire(cb, ctx, t_event, tiles, AKO_EVENT_WAVELET_END);
			data = stream;
			data_stride = G.stream_stride[g];
		}
		if (fill)
			continue;

		uint64_t* bits_at = d_bits + G.T.bits_base[g] + (uint64_t)k0 * n;
		uint8_t* blocks_at = (uint8_t*)blocks + G.T.group_base[g] + (uint64_t)k0 * n * G.T.group_stride[g];
		fire(cb, ctx, t_event, tiles, AKO_EVENT_COMPRESSION_START);
		if (st == AKO_OK && size_only)
			st = from_dev(akod_kagari_bits(ctx->dev, G.data[g] / 2, data, data_stride, bits_at, (uint32_t)members));
		else if (st == AKO_OK && s.compression != AKO_COMPRESSION_NONE)
			st = from_dev(akod_kagari_encode(ctx->dev, G.data[g] / 2, data, data_stride, blocks_at, G.T.group_stride[g],
			                                 G.T.group_cap[g], bits_at, (uint32_t)members));
		else if (st == AKO_OK)
		{
			/* no compression: the block is the raw int16 data (encode.c:151-153) */
			st = from_dev(akod_copy_strided(ctx->dev, blocks_at, G.T.group_stride[g], data, data_stride * 2, G.data[g], members));
			if (st == AKO_OK)
				st = from_dev(akod_fill_words(ctx->dev, bits_at, (uint64_t)G.data[g] * 8, members));
		}
		fire(cb, ctx, t_event, tiles, AKO_EVENT_COMPRESSION_END);
	}
	if (st != AKO_OK)
		goto done;
	if (fill)
	{
		done = n;
		goto done;
	}

	/* container assembly on the device, then one small read-back of sizes */
	if (size_only)
		st = from_dev(akod_fill_words(ctx->dev, d_total, 1, n)); /* the host adds the sizes up below */
	else if ((st = from_dev(akod_assemble(ctx->dev, head, &G.T, blocks, d_bits, s.compression != AKO_COMPRESSION_NONE, d_out,
	                                      out_stride, d_total))) != AKO_OK)
		goto done;

	{
		const size_t words = G.bits_slots + n;
		uint64_t* all = malloc(sizeof(uint64_t) * words);
		if (all == NULL)
		{
			st = AKO_NO_ENOUGH_MEMORY;
			goto done;
		}
		st = download_words(ctx, all, small, words);
		for (size_t i = 0; i < n && st == AKO_OK; i++)
		{
			uint64_t total = 16;
			for (int g = 0; g < 4; g++)
				for (size_t k = 0; k < G.count[g]; k++)
				{
					/* akoKagariEncode succeeds iff its bytes are < the capacity it was given, which is the tile's
					 * stream size minus the block head (compression.c:40-49; kagari.c:65-68, :93-107). A failure
					 * makes akoEncodeExt return AKO_ERROR (encode.c:159-164). */
					const uint64_t bytes = (all[G.T.bits_base[g] + k * n + i] + 7) / 8;
					if (s.compression != AKO_COMPRESSION_NONE && (bytes == 0 || bytes >= G.data[g] - 4))
						st = AKO_ERROR;
					total += 4 + bytes; /* head + per tile [u32 block_size][Kagari bytes] (encode.c:170-182) */
				}
			if (st == AKO_OK && all[G.bits_slots + i] == 0)
				st = AKO_ERROR;
			if (st == AKO_OK && size_only)
				all[G.bits_slots + i] = total;
			if (st == AKO_OK)
			{
				if (out_sizes != NULL)
					out_sizes[i] = (size_t)all[G.bits_slots + i];
				done = i + 1;
			}
		}
		free(all);
	}

done:
	if (out_status != NULL)
		*out_status = st;
	return done;
}

static size_t encode_core(akoB200Context* ctx, const struct akoCallbacks* cb, const struct akoSettings* s_in,
                          size_t channels, size_t w, size_t h, size_t n, const uint8_t* d_in, size_t in_stride,
                          uint8_t* d_out, size_t out_stride, size_t out_capacity, size_t* out_sizes,
                          enum akoStatus* out_status)
{
	return encode_core_ex(ctx, cb, s_in, channels, w, h, n, d_in, in_stride, d_out, out_stride, out_capacity, out_sizes,
	                      out_status, NULL);
}

AKO_API size_t akoB200EncodeBatchDevice(akoB200Context* ctx, const struct akoSettings* s, size_t channels, size_t w,
                                        size_t h, size_t n_images, const void* d_in, size_t in_stride, void* d_out,
                                        size_t out_stride, size_t* out_sizes, enum akoStatus* out_status)
{
	if (n_images == 0)
	{
		if (out_status != NULL)
			*out_status = AKO_OK;
		return 0;
	}
	/* batch members ride in a grid dimension: larger batches go through in pieces */
	size_t done = 0;
	enum akoStatus st = AKO_OK;
	while (done < n_images && st == AKO_OK)
	{
		const size_t m = (n_images - done < MEMBERS_PER_PASS) ? n_images - done : MEMBERS_PER_PASS;
		const size_t ok = encode_core(ctx, NULL, s, channels, w, h, m, (const uint8_t*)d_in + in_stride * done, in_stride,
		                              (uint8_t*)d_out + out_stride * done, out_stride, out_stride,
		                              (out_sizes != NULL) ? out_sizes + done : NULL, &st);
		done += ok;
		if (ok != m)
			break;
	}
	if (out_status != NULL)
		*out_status = st;
	return done;
}

AKO_API size_t akoB200EncodeDevice(akoB200Context* ctx, const struct akoSettings* s, size_t channels, size_t w, size_t h,
                                   const void* d_in, void* d_out, size_t out_capacity, enum akoStatus* out_status)
{
	size_t size = 0;
	const size_t ok = encode_core(ctx, NULL, s, channels, w, h, 1, d_in, 0, d_out, 0, out_capacity, &size, out_status);
	return ok ? size : 0;
}

/* ---- one tiled image over several GPUs: tiles are independent blocks (encode.c:115-205,
 * decode.c:113-230), so the tile rows are cut into bands, a few per device of $AKO_CUDA_DEVICES. A band is encoded or
 * decoded as the image it is (same width, same tiles_dimension: its tiles have the sizes they have in the whole
 * image, and a tile's block does not depend on where the tile lies); the host concatenates the bands' blocks in raster
 * order behind the one head, or cuts the blob at the block heads it has walked. Events are tile-ordered callbacks
 * and keep such calls on one device. */
#define BANDS_MIN_PIXELS ((size_t)8 << 20)
static size_t env_size(const char* name, size_t fallback, size_t lo, size_t hi);

struct band_job
{
	int device;
	size_t y0, bh;          /* image rows of the band */
	size_t t0, tn;          /* its tiles (raster order) */
	akoB200Context* ctx;
	enum akoStatus st;
	/* encode */
	const struct akoSettings* s;
	size_t channels, w;
	const uint8_t* in;
	uint8_t* d_out;
	size_t size;
	/* decode */
	const uint8_t* blob;
	const uint64_t *off, *bsz;
	uint8_t* image;
	pthread_t thread;
};

static int bands_wanted(const struct akoCallbacks* cb, const struct akoSettings* s, size_t w, size_t h, int devices[MAX_DEVICES])
{
	const size_t td = s->tiles_dimension;
	if (td == 0 || (cb != NULL && cb->events != NULL) || w * h < env_size("AKO_B200_BANDS_MIN_PIXELS", BANDS_MIN_PIXELS, 1, (size_t)1 << 40))
		return 0;
	/* Several bands per device: each band has its own pooled context and stream, so one band's copies overlap
	 * another band's kernels -- on a single device too (a 16384 x 16384 image is a gigabyte each way). */
	int listed[MAX_DEVICES];
	const int nd = env_devices(listed);
	const size_t per = env_size("AKO_B200_BANDS_PER_DEVICE", nd == 1 ? 4 : 2, 1, MAX_DEVICES);
	const size_t rows = (h / td) + (h % td != 0);
	size_t nb = (size_t)nd * per;
	nb = (nb > MAX_DEVICES) ? MAX_DEVICES : nb;
	nb = (nb > rows) ? rows : nb;
	if (nb < 2)
		return 0;
	for (size_t k = 0; k < nb; k++)
		devices[k] = listed[k % (size_t)nd];
	return (int)nb;
}

static void bands_cut(struct band_job* jobs, int nb, const int* devices, size_t w, size_t h, size_t td)
{
	const size_t rows = (h / td) + (h % td != 0), tiles_x = (w / td) + (w % td != 0);
	for (int k = 0; k < nb; k++)
	{
		const size_t r0 = rows * (size_t)k / (size_t)nb, r1 = rows * (size_t)(k + 1) / (size_t)nb;
		memset(&jobs[k], 0, sizeof(jobs[k]));
		jobs[k].device = devices[k];
		jobs[k].y0 = r0 * td;
		jobs[k].bh = ((r1 * td < h) ? r1 * td : h) - r0 * td;
		jobs[k].t0 = r0 * tiles_x;
		jobs[k].tn = (r1 - r0) * tiles_x;
		jobs[k].st = AKO_OK;
	}
}

static void* band_encode_worker(void* raw)
{
	struct band_job* j = raw;
	const struct akoCallbacks quiet = akoDefaultCallbacks();
	if ((j->ctx = pool_acquire_on(j->device, &j->st)) == NULL)
		return NULL;
	const size_t bytes = j->w * j->bh * j->channels;
	const size_t bound = akoB200EncodeBound(j->s, j->channels, j->w, j->bh);
	void *d_in, *d_out;
	if ((j->st = from_dev(akod_workspace(j->ctx->dev, AKOD_WS_INPUT, bytes + 64, &d_in))) != AKO_OK ||
	    (j->st = from_dev(akod_workspace(j->ctx->dev, AKOD_WS_OUTPUT, bound + 64, &d_out))) != AKO_OK ||
	    (j->st = from_dev(akod_h2d(j->ctx->dev, d_in, j->in + j->y0 * j->w * j->channels, bytes))) != AKO_OK)
		return NULL;
	if (encode_core(j->ctx, &quiet, j->s, j->channels, j->w, j->bh, 1, d_in, 0, d_out, 0, bound, &j->size, &j->st) != 1)
		j->size = 0;
	j->d_out = d_out;
	return NULL;
}

static size_t encode_bands(const struct akoCallbacks* cb, const struct akoSettings* s, size_t channels, size_t w, size_t h,
                           const void* in, const uint8_t head[16], const int* devices, int nb, void** out,
                           enum akoStatus* out_status)
{
	struct band_job jobs[MAX_DEVICES];
	enum akoStatus st = AKO_OK;
	uint8_t* blob = NULL;
	size_t total = 16;
	bands_cut(jobs, nb, devices, w, h, s->tiles_dimension);
	for (int k = 0; k < nb; k++)
	{
		jobs[k].s = s;
		jobs[k].channels = channels;
		jobs[k].w = w;
		jobs[k].in = in;
	}
	for (int k = 1; k < nb; k++)
		if (pthread_create(&jobs[k].thread, NULL, band_encode_worker, &jobs[k]) != 0)
		{
			jobs[k].st = AKO_ERROR;
			jobs[k].thread = 0;
		}
	band_encode_worker(&jobs[0]);
	for (int k = 1; k < nb; k++)
		if (jobs[k].thread != 0)
			pthread_join(jobs[k].thread, NULL);
	for (int k = 0; k < nb && st == AKO_OK; k++)
	{
		st = jobs[k].st;
		if (st == AKO_OK && jobs[k].size <= 16)
			st = AKO_ERROR;
		total += jobs[k].size - 16;
	}
	if (st == AKO_OK && (blob = cb->malloc(total)) == NULL)
		st = AKO_NO_ENOUGH_MEMORY;
	if (st == AKO_OK)
	{
		/* the bands' blocks, in raster order, behind the head of the whole image */
		size_t at = 16;
		memcpy(blob, head, 16);
		for (int k = 0; k < nb && st == AKO_OK; k++)
		{
			st = from_dev(akod_d2h(jobs[k].ctx->dev, blob + at, jobs[k].d_out + 16, jobs[k].size - 16));
			at += jobs[k].size - 16;
		}
		for (int k = 0; k < nb; k++)
		{
			const enum akoStatus w_st = from_dev(akod_sync(jobs[k].ctx->dev));
			if (st == AKO_OK)
				st = w_st;
		}
	}
	for (int k = 0; k < nb; k++)
		if (jobs[k].ctx != NULL)
			pool_release(jobs[k].ctx);
	if (st != AKO_OK)
	{
		if (blob != NULL)
			cb->free(blob);
		blob = NULL;
		total = 0;
	}
	if (out != NULL)
		*out = blob;
	else if (blob != NULL)
		cb->free(blob);
	if (out_status != NULL)
		*out_status = st;
	return total;
}

static enum akoStatus decode_core(akoB200Context* ctx, const struct akoCallbacks* cb, const struct akoSettings* s,
                                  size_t channels, size_t w, size_t h, size_t n, const uint8_t* d_in, size_t in_stride,
                                  const uint64_t* blk_off, const uint64_t* blk_size, uint8_t* d_out, size_t out_stride,
                                  size_t* done_out);

static void* band_decode_worker(void* raw)
{
	struct band_job* j = raw;
	const struct akoCallbacks quiet = akoDefaultCallbacks();
	uint64_t* off = NULL;
	if ((j->ctx = pool_acquire_on(j->device, &j->st)) == NULL)
		return NULL;
	/* the bytes of the band's blocks, block heads included */
	const size_t head_bytes = (j->s->compression != AKO_COMPRESSION_NONE) ? 4 : 0;
	const size_t first = (size_t)j->off[j->t0] - head_bytes;
	const size_t last = (size_t)(j->off[j->t0 + j->tn - 1] + j->bsz[j->t0 + j->tn - 1]);
	const size_t image_bytes = j->w * j->bh * j->channels;
	void *d_in, *d_out;
	size_t ok = 0;
	if ((off = malloc(sizeof(uint64_t) * j->tn)) == NULL)
	{
		j->st = AKO_NO_ENOUGH_MEMORY;
		return NULL;
	}
	for (size_t t = 0; t < j->tn; t++)
		off[t] = j->off[j->t0 + t] - first;
	if ((j->st = from_dev(akod_workspace(j->ctx->dev, AKOD_WS_INPUT, last - first + 64, &d_in))) == AKO_OK &&
	    (j->st = from_dev(akod_workspace(j->ctx->dev, AKOD_WS_OUTPUT, image_bytes + 64, &d_out))) == AKO_OK &&
	    (j->st = from_dev(akod_h2d(j->ctx->dev, d_in, j->blob + first, last - first))) == AKO_OK &&
	    (j->st = decode_core(j->ctx, &quiet, j->s, j->channels, j->w, j->bh, 1, d_in, 0, off, j->bsz + j->t0, d_out, 0, &ok)) == AKO_OK &&
	    (j->st = from_dev(akod_d2h(j->ctx->dev, j->image + j->y0 * j->w * j->channels, d_out, image_bytes))) == AKO_OK)
		j->st = from_dev(akod_sync(j->ctx->dev));
	free(off);
	return NULL;
}

static enum akoStatus decode_bands(const struct akoSettings* s, size_t channels, size_t w, size_t h, const uint8_t* blob,
                                   const uint64_t* off, const uint64_t* bsz, uint8_t* image, const int* devices, int nb)
{
	struct band_job jobs[MAX_DEVICES];
	enum akoStatus st = AKO_OK;
	bands_cut(jobs, nb, devices, w, h, s->tiles_dimension);
	for (int k = 0; k < nb; k++)
	{
		jobs[k].s = s;
		jobs[k].channels = channels;
		jobs[k].w = w;
		jobs[k].blob = blob;
		jobs[k].off = off;
		jobs[k].bsz = bsz;
		jobs[k].image = image;
	}
	for (int k = 1; k < nb; k++)
		if (pthread_create(&jobs[k].thread, NULL, band_decode_worker, &jobs[k]) != 0)
		{
			jobs[k].st = AKO_ERROR;
			jobs[k].thread = 0;
		}
	band_decode_worker(&jobs[0]);
	for (int k = 1; k < nb; k++)
		if (jobs[k].thread != 0)
			pthread_join(jobs[k].thread, NULL);
	for (int k = 0; k < nb; k++)
	{
		if (st == AKO_OK)
			st = jobs[k].st;
		if (jobs[k].ctx != NULL)
			pool_release(jobs[k].ctx);
	}
	return st;
}

AKO_API size_t akoEncodeExt(const struct akoCallbacks* c, const struct akoSettings* s, size_t channels, size_t w,
                            size_t h, const void* in, void** out, enum akoStatus* out_status)
{
	enum akoStatus st = AKO_OK;
	const struct akoCallbacks cb = (c != NULL) ? *c : akoDefaultCallbacks();
	struct akoSettings checked = (s != NULL) ? *s : akoDefaultSettings();
	akoB200Context* ctx = NULL;
	uint8_t* blob = NULL;
	size_t size = 0;

	/* check order of encode.c:53-82: callbacks, input, (allocation), header validation */
	if (cb.malloc == NULL || cb.realloc == NULL || cb.free == NULL)
	{
		st = AKO_INVALID_CALLBACKS;
		goto done;
	}
	if (in == NULL)
	{
		st = AKO_INVALID_INPUT;
		goto done;
	}
	{
		/* validate before touching the device so that bad arguments never cost a context */
		struct akoSettings v = checked;
		uint8_t head[16];
		int devices[MAX_DEVICES];
		if (v.color == AKO_COLOR_YCOCG && (v.quantization > 0 || v.gate > 0))
			v.color = AKO_COLOR_YCOCG_Q;
		if ((st = head_write(channels, w, h, &v, head)) != AKO_OK)
			goto done;
		const int nb = bands_wanted(&cb, &checked, w, h, devices);
		if (nb > 1) /* a tiled image and several GPUs: one band of tile rows per device */
			return encode_bands(&cb, &checked, channels, w, h, in, head, devices, nb, out, out_status);
	}
	if ((ctx = pool_acquire(&st)) == NULL)
		goto done;

	{
		const size_t image_bytes = w * h * channels;
		const size_t bound = akoB200EncodeBound(&checked, channels, w, h);
		void *d_in, *d_out;
		if ((st = from_dev(akod_workspace(ctx->dev, AKOD_WS_INPUT, image_bytes + 64, &d_in))) != AKO_OK ||
		    (st = from_dev(akod_workspace(ctx->dev, AKOD_WS_OUTPUT, bound + 64, &d_out))) != AKO_OK ||
		    (st = from_dev(akod_h2d(ctx->dev, d_in, in, image_bytes))) != AKO_OK)
			goto done;

		if (encode_core(ctx, &cb, &checked, channels, w, h, 1, d_in, 0, d_out, 0, bound, &size, &st) != 1)
		{
			size = 0;
			goto done;
		}
		if ((blob = cb.malloc(size)) == NULL)
		{
			st = AKO_NO_ENOUGH_MEMORY;
			size = 0;
			goto done;
		}
		if ((st = from_dev(akod_d2h(ctx->dev, blob, d_out, size))) != AKO_OK || (st = from_dev(akod_sync(ctx->dev))) != AKO_OK)
		{
			cb.free(blob);
			blob = NULL;
			size = 0;
			goto done;
		}
	}

	if (out != NULL)
		*out = blob;
	else
		cb.free(blob); /* size query only, as in encode.c:214-217 */

done:
	if (ctx != NULL)
		pool_release(ctx);
	if (out_status != NULL)
		*out_status = st;
	return size;
}

/* ---- multi-pass ratio search (tools/akoenc.cpp:111-213, EncodePass) ---- */

struct ratio_search
{
	akoB200Context* ctx;
	struct akoSettings s; /* the caller's settings; quantization is replaced per pass */
	size_t channels, w, h;
	const uint8_t* d_in;
	struct search_cache cache;
	int cache_color; /* resolved colour the cache holds, -1: empty */
	size_t passes;
	enum akoStatus st; /* status of the last pass */
};

/* One pass of the search: the size akoEncodeExt would return for quantisation q (0 when it would fail). Only the
 * first pass per colour model runs the colour transform and the wavelet; the rest re-quantise the cached stream. */
static size_t ratio_probe(struct ratio_search* r, int q)
{
	struct akoSettings s = r->s;
	s.quantization = q;
	struct akoSettings resolved = s;
	resolve_color(&resolved);
	r->passes++;

	if (r->cache_color != (int)resolved.color)
	{
		struct akoSettings raw = resolved;
		raw.quantization = 0;
		raw.gate = 0;
		struct search_cache fill = r->cache;
		fill.mode = SC_FILL;
		fill.size_only = 0;
		if (encode_core_ex(r->ctx, NULL, &raw, r->channels, r->w, r->h, 1, r->d_in, 0, NULL, 0, 0, NULL, &r->st, &fill) != 1)
			return 0;
		r->cache_color = (int)resolved.color;
	}

	size_t size = 0;
	struct search_cache use = r->cache;
	use.mode = SC_USE;
	use.size_only = 1;
	if (encode_core_ex(r->ctx, NULL, &s, r->channels, r->w, r->h, 1, r->d_in, 0, NULL, 0, 0, &size, &r->st, &use) != 1)
		return 0;
	return size;
}

static size_t absdiff(size_t a, size_t b)
{
	return (a > b) ? a - b : b - a;
}

AKO_API size_t akoB200EncodeRatio(const struct akoCallbacks* c, const struct akoSettings* s_in, int ratio,
                                  size_t channels, size_t w, size_t h, const void* in, void** out,
                                  int* out_quantization, size_t* out_passes, enum akoStatus* out_status)
{
	enum akoStatus st = AKO_OK;
	const struct akoCallbacks cb = (c != NULL) ? *c : akoDefaultCallbacks();
	struct akoSettings s = (s_in != NULL) ? *s_in : akoDefaultSettings();
	struct ratio_search r;
	uint8_t* blob = NULL;
	size_t size = 0;
	memset(&r, 0, sizeof(r));

	/* akoenc.cpp:115-128: one plain pass */
	if (ratio <= 1 || s.wavelet == AKO_WAVELET_NONE || s.compression == AKO_COMPRESSION_NONE)
	{
		if (ratio == 1 && s.wavelet != AKO_WAVELET_NONE && s.compression != AKO_COMPRESSION_NONE)
		{
			s.quantization = 0;
			s.gate = 0;
		}
		if (out_quantization != NULL)
			*out_quantization = s.quantization;
		if (out_passes != NULL)
			*out_passes = 1;
		return akoEncodeExt(c, &s, channels, w, h, in, out, out_status);
	}

	if (cb.malloc == NULL || cb.realloc == NULL || cb.free == NULL)
	{
		st = AKO_INVALID_CALLBACKS;
		goto done;
	}
	if (in == NULL)
	{
		st = AKO_INVALID_INPUT;
		goto done;
	}
	{
		struct akoSettings v = s;
		uint8_t head[16];
		resolve_color(&v);
		if ((st = head_write(channels, w, h, &v, head)) != AKO_OK)
			goto done;
		if (channels == 0)
		{
			st = AKO_INVALID_CHANNELS_NO;
			goto done;
		}
	}
	if ((r.ctx = pool_acquire(&st)) == NULL)
		goto done;

	{
		const size_t image_bytes = w * h * channels;
		const size_t td = s.tiles_dimension;
		const size_t tiles = tiles_count(w, h, td);
		size_t cache_elems = 0;
		{
			size_t tx = 0, ty = 0;
			for (size_t t = 0; t < tiles; t++)
			{
				cache_elems += align_up(tile_data_size(tile_dimension(tx, w, td), tile_dimension(ty, h, td)) * channels / 2, 8);
				tx += td;
				if (tx >= w)
				{
					tx = 0;
					ty += td;
				}
			}
		}
		void *d_in, *d_raw;
		if ((st = from_dev(akod_workspace(r.ctx->dev, AKOD_WS_INPUT, image_bytes + 64, &d_in))) != AKO_OK ||
		    (st = from_dev(akod_workspace(r.ctx->dev, AKOD_WS_SEARCH, cache_elems * 2 + 64, &d_raw))) != AKO_OK ||
		    (st = from_dev(akod_h2d(r.ctx->dev, d_in, in, image_bytes))) != AKO_OK)
			goto done;
		r.s = s;
		r.channels = channels;
		r.w = w;
		r.h = h;
		r.d_in = d_in;
		r.cache.d_raw = d_raw;
		r.cache_color = -1;

		/* akoenc.cpp:130-192, sizes come from the Kagari length scan */
		const size_t target_size = (w * h * channels) / (size_t)ratio;
		const size_t error_margin = (target_size * 4) / 100;
		size_t ceil_size = ratio_probe(&r, 0);
		int q = 1, floor_q = 0, ceil_q = 0, last_q = 0;
		size_t floor_size = ceil_size;
		do
		{
			q *= 4;
			ceil_size = floor_size;
			ceil_q = floor_q;
			floor_size = ratio_probe(&r, q);
			floor_q = q;
			last_q = q;
		} while (floor_size > target_size && q <= (1 << 28)); /* the tool's q overflows here; every q this large
		                                                           quantises alike (quantization.c:86-96) */
		size_t last_size = floor_size;
		while (absdiff(floor_size, ceil_size) > error_margin && abs(floor_q - ceil_q) > 1)
		{
			q = (ceil_q + floor_q) / 2;
			last_size = ratio_probe(&r, q);
			last_q = q;
			if (last_size > target_size)
			{
				ceil_size = last_size;
				ceil_q = q;
			}
			else
			{
				floor_size = last_size;
				floor_q = q;
			}
		}

		/* akoenc.cpp:194-210: the tool keeps the LAST pass's blob whenever its size equals the chosen bound's */
		const int pick_floor = absdiff(floor_size, target_size) < absdiff(ceil_size, target_size);
		const size_t chosen_size = pick_floor ? floor_size : ceil_size;
		int final_q = pick_floor ? floor_q : ceil_q;
		if (last_size == chosen_size)
			final_q = last_q;
		else
			r.passes++; /* the tool encodes once more */
		if (chosen_size == 0)
		{
			st = (r.st != AKO_OK) ? r.st : AKO_ERROR;
			goto done;
		}

		/* the blob itself: re-quantise the cached stream once more, pack, assemble */
		struct akoSettings fs = s;
		fs.quantization = final_q;
		struct akoSettings resolved = fs;
		resolve_color(&resolved);
		const size_t bound = akoB200EncodeBound(&fs, channels, w, h);
		void* d_out;
		if ((st = from_dev(akod_workspace(r.ctx->dev, AKOD_WS_OUTPUT, bound + 64, &d_out))) != AKO_OK)
			goto done;
		struct search_cache use = r.cache;
		use.mode = SC_USE;
		use.size_only = 0;
		const struct search_cache* how = (r.cache_color == (int)resolved.color) ? &use : NULL;
		if (encode_core_ex(r.ctx, &cb, &fs, channels, w, h, 1, d_in, 0, d_out, 0, bound, &size, &st, how) != 1)
		{
			size = 0;
			goto done;
		}
		if ((blob = cb.malloc(size)) == NULL)
		{
			st = AKO_NO_ENOUGH_MEMORY;
			size = 0;
			goto done;
		}
		if ((st = from_dev(akod_d2h(r.ctx->dev, blob, d_out, size))) != AKO_OK ||
		    (st = from_dev(akod_sync(r.ctx->dev))) != AKO_OK)
		{
			cb.free(blob);
			blob = NULL;
			size = 0;
			goto done;
		}
		if (out_quantization != NULL)
			*out_quantization = final_q;
	}

	if (out != NULL)
		*out = blob;
	else
		cb.free(blob);

done:
	if (out_passes != NULL)
		*out_passes = r.passes;
	if (r.ctx != NULL)
		pool_release(r.ctx);
	if (out_status != NULL)
		*out_status = st;
	return size;
}

/* ---- decode ---- */

/* n same-shape blobs. Block offsets/sizes per image and tile are host arrays [n][tiles]. */
static enum akoStatus decode_core(akoB200Context* ctx, const struct akoCallbacks* cb, const struct akoSettings* s,
                                  size_t channels, size_t w, size_t h, size_t n, const uint8_t* d_in, size_t in_stride,
                                  const uint64_t* blk_off, const uint64_t* blk_size, uint8_t* d_out, size_t out_stride,
                                  size_t* done_out)
{
	enum akoStatus st = AKO_OK;
	const size_t td = s->tiles_dimension;
	struct tile_groups G;
	make_groups(&G, s, channels, w, h, n);
	const size_t tiles = G.tiles;
	const int tile_by_tile = (cb != NULL && cb->events != NULL);
	uint64_t* off_abs = NULL;
	*done_out = 0;
	if (n > MEMBERS_PER_PASS)
		return AKO_ERROR; /* the batch entry points split larger batches */

	void *planes, *scratch, *stream, *small;
	{
		size_t need_planes = 0, need_scratch = 0, need_stream = 0;
		for (int g = 0; g < 4; g++)
			if (G.count[g] != 0)
			{
				const size_t members = group_pass_tiles(&G, g, n, tile_by_tile) * n;
				need_planes = (G.planes_stride[g] * members > need_planes) ? G.planes_stride[g] * members : need_planes;
				need_scratch = (G.scratch_stride[g] * members > need_scratch) ? G.scratch_stride[g] * members : need_scratch;
				need_stream = (G.stream_stride[g] * members > need_stream) ? G.stream_stride[g] * members : need_stream;
			}
		if ((st = from_dev(akod_workspace(ctx->dev, AKOD_WS_PLANES, need_planes * 2 + 64, &planes))) != AKO_OK ||
		    (st = from_dev(akod_workspace(ctx->dev, AKOD_WS_SCRATCH, need_scratch * 2 + 64, &scratch))) != AKO_OK ||
		    (st = from_dev(akod_workspace(ctx->dev, AKOD_WS_STREAM, need_stream * 2 + 64, &stream))) != AKO_OK ||
		    (st = from_dev(akod_workspace(ctx->dev, AKOD_WS_SMALL, sizeof(uint64_t) * G.bits_slots * 3 + 64, &small))) != AKO_OK)
			return st;
	}
	const size_t slots = G.bits_slots;
	uint64_t* d_result = small;        /* all three laid out as akodTiles says */
	uint64_t* d_off = d_result + slots; /* absolute byte offsets into d_in */
	uint64_t* d_size = d_off + slots;

	off_abs = malloc(sizeof(uint64_t) * slots * 3);
	if (off_abs == NULL)
		return AKO_NO_ENOUGH_MEMORY;
	uint64_t* size_abs = off_abs + slots;
	uint64_t* results = size_abs + slots;
	uint64_t largest[4] = {0, 0, 0, 0};
	for (size_t t = 0; t < tiles; t++)
	{
		uint32_t g, k;
		akod_tile_locate(&G.T, (uint32_t)t, &g, &k);
		for (size_t i = 0; i < n; i++)
		{
			const size_t at = (size_t)G.T.bits_base[g] + (size_t)k * n + i;
			off_abs[at] = in_stride * i + blk_off[tiles * i + t];
			size_abs[at] = blk_size[tiles * i + t];
			largest[g] = (size_abs[at] > largest[g]) ? size_abs[at] : largest[g];
		}
	}
	if (slots * 2 <= MAILBOX_WORDS)
		st = upload_words(ctx, d_off, off_abs, slots * 2);
	else
	{
		st = from_dev(akod_h2d(ctx->dev, d_off, off_abs, sizeof(uint64_t) * slots * 2));
		if (st == AKO_OK)
			st = from_dev(akod_sync(ctx->dev)); /* pageable source: off_abs may be reused only after the copy */
	}
	if (st != AKO_OK)
	{
		free(off_abs);
		return st;
	}

	for (size_t item = 0; st == AKO_OK; item++)
	{
		uint32_t g = 0, k0 = 0;
		size_t kc = 0, t_event = 0;
		if (tile_by_tile)
		{
			if (item >= tiles)
				break;
			akod_tile_locate(&G.T, (uint32_t)item, &g, &k0);
			kc = 1;
			t_event = item;
		}
		else
		{
			size_t left = item;
			int found = 0;
			for (int gg = 0; gg < 4 && !found; gg++)
			{
				if (G.count[gg] == 0)
					continue;
				const size_t per = group_pass_tiles(&G, gg, n, 0);
				const size_t passes = (G.count[gg] + per - 1) / per;
				if (left < passes)
				{
					g = (uint32_t)gg;
					k0 = (uint32_t)(left * per);
					kc = (G.count[gg] - k0 < per) ? G.count[gg] - k0 : per;
					found = 1;
				}
				else
					left -= passes;
			}
			if (!found)
				break;
		}
		const size_t tw = G.tw[g], th = G.th[g];
		const size_t members = kc * n;
		const size_t at = (size_t)G.T.bits_base[g] + (size_t)k0 * n;
		akodBatch batch;
		group_batch(&batch, &G, (int)g, k0, kc, n, out_stride, td); /* format_inverse reads in_stride as the u8 image stride */
		int16_t* target = (s->wavelet != AKO_WAVELET_NONE) ? stream : planes;
		const uint64_t target_stride = (s->wavelet != AKO_WAVELET_NONE) ? G.stream_stride[g] : G.planes_stride[g];

		fire(cb, ctx, t_event, tiles, AKO_EVENT_COMPRESSION_START);
		if (s->compression != AKO_COMPRESSION_NONE)
			st = from_dev(akod_kagari_decode(ctx->dev, G.data[g] / 2, d_in, d_off + at, d_size + at, largest[g], target,
			                                 target_stride, d_result + at, (uint32_t)members));
		else
			st = from_dev(akod_gather(ctx->dev, target, target_stride * 2, d_in, d_off + at, G.data[g], members));
		fire(cb, ctx, t_event, tiles, AKO_EVENT_COMPRESSION_END);

		if (st == AKO_OK && s->wavelet != AKO_WAVELET_NONE)
		{
			fire(cb, ctx, t_event, tiles, AKO_EVENT_WAVELET_START);
			st = from_dev(akod_unlift(ctx->dev, get_plan(ctx, s, channels, tw, th), stream, planes, scratch, &batch));
			fire(cb, ctx, t_event, tiles, AKO_EVENT_WAVELET_END);
		}
		if (st == AKO_OK)
		{
			fire(cb, ctx, t_event, tiles, AKO_EVENT_FORMAT_START);
			st = from_dev(akod_format_inverse(ctx->dev, (int)s->color, (uint32_t)channels, (uint32_t)tw, (uint32_t)th, w,
			                                  planes, d_out, &batch));
			fire(cb, ctx, t_event, tiles, AKO_EVENT_FORMAT_END);
		}
	}

	if (st == AKO_OK && s->compression != AKO_COMPRESSION_NONE)
	{
		/* compression.c:69-70, decode.c:152-156: every block must have been consumed exactly */
		size_t first_bad = n;
		if (slots <= MAILBOX_WORDS)
			st = download_words(ctx, results, d_result, slots);
		else
		{
			st = from_dev(akod_d2h(ctx->dev, results, d_result, sizeof(uint64_t) * slots));
			if (st == AKO_OK)
				st = from_dev(akod_sync(ctx->dev));
		}
		/* slot = bits_base[g] + k * n + i: the image of a slot is (slot - bits_base[g]) % n */
		for (int g = 0; g < 4 && st == AKO_OK; g++)
			for (size_t j = 0; j < G.count[g] * n; j++)
			{
				const size_t at = (size_t)G.T.bits_base[g] + j;
				if ((results[at] == 0 || results[at] != size_abs[at]) && j % n < first_bad)
					first_bad = j % n;
			}
		if (st == AKO_OK)
		{
			*done_out = first_bad;
			if (first_bad != n)
				st = AKO_BROKEN_INPUT;
		}
	}
	else if (st == AKO_OK)
	{
		st = from_dev(akod_sync(ctx->dev));
		if (st == AKO_OK)
			*done_out = n;
	}
	free(off_abs);
	return st;
}

/* walks the block heads of one blob on the host; returns AKO_BROKEN_INPUT if it runs past the end */
static enum akoStatus walk_blocks_host(const uint8_t* blob, size_t input_size, const struct akoSettings* s,
                                       size_t channels, size_t w, size_t h, uint64_t* off, uint64_t* size)
{
	const size_t td = s->tiles_dimension;
	const size_t tiles = tiles_count(w, h, td);
	size_t pos = 16, tx = 0, ty = 0;
	for (size_t t = 0; t < tiles; t++)
	{
		const size_t tw = tile_dimension(tx, w, td), th = tile_dimension(ty, h, td);
		if (s->compression != AKO_COMPRESSION_NONE)
		{
			if (pos + 4 > input_size)
				return AKO_BROKEN_INPUT;
			const size_t block = load_le32(blob + pos);
			if (block == 0 || pos + 4 + block > input_size)
				return AKO_BROKEN_INPUT;
			off[t] = pos + 4;
			size[t] = block;
			pos += 4 + block;
		}
		else
		{
			const size_t data = (s->wavelet != AKO_WAVELET_NONE) ? tile_data_size(tw, th) * channels : tw * th * channels * 2;
			if (pos + data > input_size) /* decode.c:163-167 */
				return AKO_BROKEN_INPUT;
			off[t] = pos;
			size[t] = data;
			pos += data;
		}
		tx += td;
		if (tx >= w)
		{
			tx = 0;
			ty += td;
		}
	}
	return AKO_OK;
}

AKO_API uint8_t* akoDecodeExt(const struct akoCallbacks* c, size_t input_size, const void* in, struct akoSettings* out_s,
                              size_t* out_channels, size_t* out_w, size_t* out_h, enum akoStatus* out_status)
{
	enum akoStatus st = AKO_OK;
	const struct akoCallbacks cb = (c != NULL) ? *c : akoDefaultCallbacks();
	struct akoSettings s;
	size_t channels = 0, w = 0, h = 0;
	akoB200Context* ctx = NULL;
	uint8_t* image = NULL;
	uint64_t* blk = NULL;
	memset(&s, 0, sizeof(s));

	if (cb.malloc == NULL || cb.realloc == NULL || cb.free == NULL)
	{
		st = AKO_INVALID_CALLBACKS;
		goto done;
	}
	if (in == NULL)
	{
		st = AKO_INVALID_INPUT;
		goto done;
	}
	if (input_size < 16) /* the reference's own check is a tautology (decode.c:71); this one is real */
	{
		st = AKO_BROKEN_INPUT;
		goto done;
	}
	if ((st = head_read(in, &channels, &w, &h, &s)) != AKO_OK)
		goto done;
	if (s.wavelet == AKO_WAVELET_NONE && s.compression != AKO_COMPRESSION_NONE)
	{
		st = AKO_ERROR; /* see encode_core */
		goto done;
	}

	const size_t tiles = tiles_count(w, h, s.tiles_dimension);
	if ((blk = malloc(sizeof(uint64_t) * tiles * 2)) == NULL)
	{
		st = AKO_NO_ENOUGH_MEMORY;
		goto done;
	}
	if ((st = walk_blocks_host(in, input_size, &s, channels, w, h, blk, blk + tiles)) != AKO_OK)
		goto done;
	{
		int devices[MAX_DEVICES];
		const int nb = bands_wanted(&cb, &s, w, h, devices);
		if (nb > 1) /* a tiled image and several GPUs: one band of tile rows per device */
		{
			if ((image = cb.malloc(w * h * channels)) == NULL)
				st = AKO_NO_ENOUGH_MEMORY;
			else if ((st = decode_bands(&s, channels, w, h, in, blk, blk + tiles, image, devices, nb)) != AKO_OK)
			{
				cb.free(image);
				image = NULL;
			}
			if (st == AKO_OK)
			{
				if (out_s != NULL)
					*out_s = s;
				if (out_channels != NULL)
					*out_channels = channels;
				if (out_w != NULL)
					*out_w = w;
				if (out_h != NULL)
					*out_h = h;
			}
			goto done;
		}
	}
	if ((ctx = pool_acquire(&st)) == NULL)
		goto done;

	{
		const size_t image_bytes = w * h * channels;
		void *d_in, *d_out;
		size_t ok = 0;
		if ((st = from_dev(akod_workspace(ctx->dev, AKOD_WS_INPUT, input_size + 64, &d_in))) != AKO_OK ||
		    (st = from_dev(akod_workspace(ctx->dev, AKOD_WS_OUTPUT, image_bytes + 64, &d_out))) != AKO_OK ||
		    (st = from_dev(akod_h2d(ctx->dev, d_in, in, input_size))) != AKO_OK)
			goto done;
		if ((st = decode_core(ctx, &cb, &s, channels, w, h, 1, d_in, 0, blk, blk + tiles, d_out, 0, &ok)) != AKO_OK)
			goto done;
		if ((image = cb.malloc(image_bytes)) == NULL)
		{
			st = AKO_NO_ENOUGH_MEMORY;
			goto done;
		}
		if ((st = from_dev(akod_d2h(ctx->dev, image, d_out, image_bytes))) != AKO_OK ||
		    (st = from_dev(akod_sync(ctx->dev))) != AKO_OK)
		{
			cb.free(image);
			image = NULL;
			goto done;
		}
	}

	if (out_s != NULL)
		*out_s = s;
	if (out_channels != NULL)
		*out_channels = channels;
	if (out_w != NULL)
		*out_w = w;
	if (out_h != NULL)
		*out_h = h;

done:
	free(blk);
	if (ctx != NULL)
		pool_release(ctx);
	if (out_status != NULL)
		*out_status = st;
	return image;
}

/* device-resident blobs: the block heads are walked on the device, then read back */
static enum akoStatus walk_blocks_device(akoB200Context* ctx, const uint8_t* d_blob, size_t input_size,
                                         const struct akoSettings* s, size_t channels, size_t w, size_t h, uint64_t* off,
                                         uint64_t* size)
{
	const size_t tiles = tiles_count(w, h, s->tiles_dimension);
	if (s->compression == AKO_COMPRESSION_NONE)
	{
		/* sizes are pure geometry */
		uint8_t fake[4] = {0, 0, 0, 0};
		(void)fake;
		size_t pos = 16, tx = 0, ty = 0;
		const size_t td = s->tiles_dimension;
		for (size_t t = 0; t < tiles; t++)
		{
			const size_t tw = tile_dimension(tx, w, td), th = tile_dimension(ty, h, td);
			const size_t data = (s->wavelet != AKO_WAVELET_NONE) ? tile_data_size(tw, th) * channels : tw * th * channels * 2;
			if (pos + data > input_size)
				return AKO_BROKEN_INPUT;
			off[t] = pos;
			size[t] = data;
			pos += data;
			tx += td;
			if (tx >= w)
			{
				tx = 0;
				ty += td;
			}
		}
		return AKO_OK;
	}

	enum akoStatus st;
	void* small;
	uint64_t* mail = akod_mailbox(ctx->dev);
	if (tiles * 2 * sizeof(uint64_t) > (1 << 16))
		return AKO_ERROR; /* more than 4096 tiles per device-resident blob: use the host-pointer API */
	if ((st = from_dev(akod_workspace(ctx->dev, AKOD_WS_SMALL, sizeof(uint64_t) * tiles * 2 + 64, &small))) != AKO_OK)
		return st;
	uint64_t* d_off = small;
	if ((st = from_dev(akod_walk_blocks(ctx->dev, d_blob, input_size, (uint32_t)tiles, d_off, d_off + tiles))) != AKO_OK ||
	    (st = from_dev(akod_d2h(ctx->dev, mail, d_off, sizeof(uint64_t) * tiles * 2))) != AKO_OK ||
	    (st = from_dev(akod_sync(ctx->dev))) != AKO_OK)
		return st;
	for (size_t t = 0; t < tiles; t++)
	{
		off[t] = mail[t];
		size[t] = mail[tiles + t];
		if (size[t] == 0)
			return AKO_BROKEN_INPUT;
	}
	return AKO_OK;
}

/* The 16-byte head of a device-resident blob and, in the same read-back, the size field of its first block (bytes
 * 16..19; *first = 0 when the blob ends before it). An untiled image has no other block: its block walk is then done
 * (first_block), which saves a lone image one kernel and one round trip. */
static enum akoStatus fetch_head(akoB200Context* ctx, const void* d_in, size_t input_size, uint8_t head[16], uint64_t* first)
{
	uint8_t* mail = akod_mailbox(ctx->dev);
	const size_t n = (input_size >= 20) ? 20 : 16;
	enum akoStatus st;
	if ((st = from_dev(akod_d2h(ctx->dev, mail, d_in, n))) != AKO_OK || (st = from_dev(akod_sync(ctx->dev))) != AKO_OK)
		return st;
	memcpy(head, mail, 16);
	*first = (n == 20) ? (uint64_t)load_le32(mail + 16) : 0;
	return AKO_OK;
}

/* what k_walk_blocks answers for the only block of an untiled blob (decode.c:150-158) */
static enum akoStatus first_block(uint64_t first, size_t input_size, uint64_t* off, uint64_t* size)
{
	if (first == 0 || 20 + first > input_size)
		return AKO_BROKEN_INPUT;
	*off = 20;
	*size = first;
	return AKO_OK;
}

AKO_API enum akoStatus akoB200DecodeDevice(akoB200Context* ctx, size_t input_size, const void* d_in, const void* head16,
                                           void* d_out, size_t out_capacity, struct akoSettings* out_s,
                                           size_t* out_channels, size_t* out_w, size_t* out_h)
{
	enum akoStatus st;
	struct akoSettings s;
	size_t channels = 0, w = 0, h = 0, ok = 0;
	uint8_t head[16];
	uint64_t* blk = NULL;
	memset(&s, 0, sizeof(s));

	if (d_in == NULL || d_out == NULL)
		return AKO_INVALID_INPUT;
	if (input_size < 16)
		return AKO_BROKEN_INPUT;
	uint64_t first = 0;
	if (head16 != NULL)
		memcpy(head, head16, 16);
	else if ((st = fetch_head(ctx, d_in, input_size, head, &first)) != AKO_OK)
		return st;
	if ((st = head_read(head, &channels, &w, &h, &s)) != AKO_OK)
		return st;
	if (s.wavelet == AKO_WAVELET_NONE && s.compression != AKO_COMPRESSION_NONE)
		return AKO_ERROR;
	if (out_capacity < w * h * channels)
		return AKO_NO_ENOUGH_MEMORY;

	const size_t tiles = tiles_count(w, h, s.tiles_dimension);
	if ((blk = malloc(sizeof(uint64_t) * tiles * 2)) == NULL)
		return AKO_NO_ENOUGH_MEMORY;
	if (head16 == NULL && tiles == 1 && s.compression != AKO_COMPRESSION_NONE)
		st = first_block(first, input_size, blk, blk + 1);
	else
		st = walk_blocks_device(ctx, d_in, input_size, &s, channels, w, h, blk, blk + tiles);
	if (st == AKO_OK)
		st = decode_core(ctx, NULL, &s, channels, w, h, 1, d_in, 0, blk, blk + tiles, d_out, 0, &ok);
	free(blk);
	if (st != AKO_OK)
		return st;

	if (out_s != NULL)
		*out_s = s;
	if (out_channels != NULL)
		*out_channels = channels;
	if (out_w != NULL)
		*out_w = w;
	if (out_h != NULL)
		*out_h = h;
	return AKO_OK;
}

AKO_API size_t akoB200DecodeBatchDevice(akoB200Context* ctx, size_t n_images, const void* d_in, size_t in_stride,
                                        const size_t* in_sizes, void* d_out, size_t out_stride,
                                        enum akoStatus* out_status)
{
	enum akoStatus st = AKO_OK;
	struct akoSettings s;
	size_t channels = 0, w = 0, h = 0, ok = 0;
	uint64_t* blk = NULL;
	memset(&s, 0, sizeof(s));

	if (n_images == 0)
		goto done;
	if (d_in == NULL || d_out == NULL || in_sizes == NULL)
	{
		st = AKO_INVALID_INPUT;
		goto done;
	}
	uint64_t first = 0;
	{
		/* header of blob 0 defines the batch's shape */
		if (in_sizes[0] < 16)
		{
			st = AKO_BROKEN_INPUT;
			goto done;
		}
		uint8_t head[16];
		if ((st = fetch_head(ctx, d_in, in_sizes[0], head, &first)) != AKO_OK)
			goto done;
		if ((st = head_read(head, &channels, &w, &h, &s)) != AKO_OK)
			goto done;
	}
	if (s.wavelet == AKO_WAVELET_NONE && s.compression != AKO_COMPRESSION_NONE)
	{
		st = AKO_ERROR;
		goto done;
	}
	if (out_stride < w * h * channels)
	{
		st = AKO_NO_ENOUGH_MEMORY;
		goto done;
	}

	const size_t tiles = tiles_count(w, h, s.tiles_dimension);
	if ((blk = malloc(sizeof(uint64_t) * tiles * 2 * n_images)) == NULL)
	{
		st = AKO_NO_ENOUGH_MEMORY;
		goto done;
	}
	uint64_t* off = blk;
	uint64_t* size = blk + tiles * n_images;
	if (s.compression == AKO_COMPRESSION_NONE)
	{
		for (size_t i = 0; i < n_images && st == AKO_OK; i++)
			st = walk_blocks_device(ctx, (const uint8_t*)d_in + in_stride * i, in_sizes[i], &s, channels, w, h,
			                        off + tiles * i, size + tiles * i);
	}
	else if (n_images == 1 && tiles == 1)
		st = first_block(first, in_sizes[0], off, size); /* one untiled image: the head's read-back held its block walk */
	else
	{
		/* one kernel walks the block heads of every blob, one read-back brings offsets and sizes home */
		void* small;
		uint64_t* sizes64 = malloc(sizeof(uint64_t) * n_images);
		if (sizes64 == NULL)
			st = AKO_NO_ENOUGH_MEMORY;
		if (st == AKO_OK)
			st = from_dev(akod_workspace(ctx->dev, AKOD_WS_SMALL, sizeof(uint64_t) * n_images * (2 * tiles + 1) + 64, &small));
		if (st == AKO_OK)
		{
			uint64_t* d_sizes = small;
			uint64_t* d_off = d_sizes + n_images;
			for (size_t i = 0; i < n_images; i++)
				sizes64[i] = in_sizes[i];
			if ((st = upload_words(ctx, d_sizes, sizes64, n_images)) == AKO_OK &&
			    (st = from_dev(akod_walk_blocks_batch(ctx->dev, d_in, in_stride, d_sizes, (uint32_t)tiles, (uint32_t)n_images,
			                                          d_off, d_off + tiles * n_images))) == AKO_OK)
				st = download_words(ctx, blk, d_off, 2 * tiles * n_images);
			for (size_t k = 0; k < tiles * n_images && st == AKO_OK; k++)
				if (size[k] == 0)
					st = AKO_BROKEN_INPUT;
		}
		free(sizes64);
	}
	while (st == AKO_OK && ok < n_images)
	{
		const size_t m = (n_images - ok < MEMBERS_PER_PASS) ? n_images - ok : MEMBERS_PER_PASS;
		size_t part = 0;
		st = decode_core(ctx, NULL, &s, channels, w, h, m, (const uint8_t*)d_in + in_stride * ok, in_stride, off + tiles * ok,
		                 size + tiles * ok, (uint8_t*)d_out + out_stride * ok, out_stride, &part);
		ok += part;
		if (part != m)
			break;
	}

done:
	free(blk);
	if (out_status != NULL)
		*out_status = st;
	return ok;
}

/* ------------------------------------------------------------------------------------------------ */
/* batches of host images / host blobs (additive; SURVEY 8b "batch entry point", 8f4 "pipelined host I/O")      */

/* The batch is cut into chunks of a few images. A chunk is one pass of the batched kernels (encode_core /
 * decode_core with n = chunk) on its own pooled context and stream, bracketed by the host<->device copies of its
 * images. A few worker threads (the caller is one of them) take chunks as they come, so the copies of one chunk
 * overlap the kernels of another and the read-backs of a third: PCIe stays busy in both directions. */
#define HOST_BATCH_CHUNK 8    /* most images per chunk ($AKO_B200_BATCH_CHUNK lowers it) */
#define HOST_BATCH_WORKERS 32 /* most worker threads ($AKO_B200_BATCH_WORKERS; default two per device) */

static size_t env_size(const char* name, size_t fallback, size_t lo, size_t hi)
{
	const char* v = getenv(name);
	if (v == NULL || v[0] == '\0')
		return fallback;
	const long x = strtol(v, NULL, 10);
	return (x < (long)lo) ? lo : (x > (long)hi) ? hi : (size_t)x;
}

struct host_batch
{
	int decode;
	struct akoCallbacks cb;
	struct akoSettings s; /* encode: the caller's; decode: what blob 0's head says */
	size_t channels, w, h, n, chunk, workers;
	int devices[MAX_DEVICES];
	int n_devices;
	const void* const* in;
	const size_t* in_sizes;
	void** out;
	size_t* out_sizes;

	pthread_mutex_t lock;
	size_t next_chunk;
	size_t first_failed; /* index of the first image that failed, n when none did */
	enum akoStatus st;   /* status of that image */
};

static void host_batch_fail(struct host_batch* b, size_t image, enum akoStatus st)
{
	pthread_mutex_lock(&b->lock);
	if (image < b->first_failed)
	{
		b->first_failed = image;
		b->st = st;
	}
	pthread_mutex_unlock(&b->lock);
}

static void host_batch_encode_chunk(struct host_batch* b, size_t first, size_t count, int device)
{
	enum akoStatus st = AKO_OK;
	akoB200Context* ctx = pool_acquire_on(device, &st);
	if (ctx == NULL)
	{
		host_batch_fail(b, first, st);
		return;
	}
	const size_t image_bytes = b->w * b->h * b->channels;
	const size_t in_stride = align_up(image_bytes, 256);
	const size_t out_stride = align_up(akoB200EncodeBound(&b->s, b->channels, b->w, b->h), 256);
	void *d_in, *d_out;
	size_t sizes[HOST_BATCH_CHUNK];
	size_t done = 0;
	if ((st = from_dev(akod_workspace(ctx->dev, AKOD_WS_INPUT, in_stride * count + 64, &d_in))) == AKO_OK)
		st = from_dev(akod_workspace(ctx->dev, AKOD_WS_OUTPUT, out_stride * count + 64, &d_out));
	for (size_t i = 0; i < count && st == AKO_OK; i++)
	{
		if (b->in[first + i] == NULL)
			st = AKO_INVALID_INPUT;
		else
			st = from_dev(akod_h2d(ctx->dev, (uint8_t*)d_in + in_stride * i, b->in[first + i], image_bytes));
	}
	if (st == AKO_OK)
		done = encode_core(ctx, NULL, &b->s, b->channels, b->w, b->h, count, d_in, in_stride, d_out, out_stride, out_stride,
		                   sizes, &st);
	for (size_t i = 0; i < done; i++)
	{
		void* blob = b->cb.malloc(sizes[i]);
		if (blob == NULL)
		{
			st = AKO_NO_ENOUGH_MEMORY;
			done = i;
			break;
		}
		b->out[first + i] = blob;
		b->out_sizes[first + i] = sizes[i];
		const enum akoStatus cst = from_dev(akod_d2h(ctx->dev, blob, (uint8_t*)d_out + out_stride * i, sizes[i]));
		if (cst != AKO_OK)
		{
			st = cst;
			done = i;
			break;
		}
	}
	{
		const enum akoStatus sst = from_dev(akod_sync(ctx->dev));
		if (sst != AKO_OK && st == AKO_OK)
		{
			st = sst;
			done = 0;
		}
	}
	if (done != count)
		host_batch_fail(b, first + done, (st != AKO_OK) ? st : AKO_ERROR);
	pool_release(ctx);
}

static void host_batch_decode_chunk(struct host_batch* b, size_t first, size_t count, int device)
{
	enum akoStatus st = AKO_OK;
	const size_t tiles = tiles_count(b->w, b->h, b->s.tiles_dimension);
	const size_t image_bytes = b->w * b->h * b->channels;
	uint64_t* blk = malloc(sizeof(uint64_t) * tiles * 2 * count);
	akoB200Context* ctx = NULL;
	size_t done = 0, usable = count, largest = 0;
	if (blk == NULL)
	{
		host_batch_fail(b, first, AKO_NO_ENOUGH_MEMORY);
		return;
	}

	/* every blob must describe the batch's shape and settings; walk its block heads on the host */
	for (size_t i = 0; i < count; i++)
	{
		struct akoSettings s;
		size_t channels = 0, w = 0, h = 0;
		const uint8_t* blob = b->in[first + i];
		const size_t size = b->in_sizes[first + i];
		enum akoStatus bst = AKO_OK;
		memset(&s, 0, sizeof(s));
		if (blob == NULL)
			bst = AKO_INVALID_INPUT;
		else if (size < 16)
			bst = AKO_BROKEN_INPUT;
		else if ((bst = head_read(blob, &channels, &w, &h, &s)) == AKO_OK)
		{
			if (channels != b->channels || w != b->w || h != b->h || memcmp(&s, &b->s, sizeof(s)) != 0)
				bst = AKO_INVALID_INPUT; /* not the shape of blob 0: decode it on its own with akoDecodeExt */
			else
				bst = walk_blocks_host(blob, size, &s, channels, w, h, blk + tiles * i, blk + tiles * (count + i));
		}
		if (bst != AKO_OK)
		{
			host_batch_fail(b, first + i, bst);
			usable = i;
			break;
		}
		largest = (size > largest) ? size : largest;
	}
	if (usable == 0 || (ctx = pool_acquire_on(device, &st)) == NULL)
	{
		if (usable != 0)
			host_batch_fail(b, first, st);
		free(blk);
		return;
	}
	if (usable != count) /* sizes of the usable prefix must be contiguous: [usable][tiles] */
		memmove(blk + tiles * usable, blk + tiles * count, sizeof(uint64_t) * tiles * usable);

	const size_t in_stride = align_up(largest, 256);
	const size_t out_stride = align_up(image_bytes, 256);
	void *d_in, *d_out;
	if ((st = from_dev(akod_workspace(ctx->dev, AKOD_WS_INPUT, in_stride * usable + 64, &d_in))) == AKO_OK)
		st = from_dev(akod_workspace(ctx->dev, AKOD_WS_OUTPUT, out_stride * usable + 64, &d_out));
	for (size_t i = 0; i < usable && st == AKO_OK; i++)
		st = from_dev(akod_h2d(ctx->dev, (uint8_t*)d_in + in_stride * i, b->in[first + i], b->in_sizes[first + i]));
	if (st == AKO_OK)
	{
		st = decode_core(ctx, NULL, &b->s, b->channels, b->w, b->h, usable, d_in, in_stride, blk, blk + tiles * usable, d_out,
		                 out_stride, &done);
		if (st == AKO_BROKEN_INPUT && done < usable)
			host_batch_fail(b, first + done, st); /* images before it are fine */
		else if (st != AKO_OK)
			done = 0;
	}
	for (size_t i = 0; i < done; i++)
	{
		uint8_t* image = b->cb.malloc(image_bytes);
		enum akoStatus cst = AKO_NO_ENOUGH_MEMORY;
		if (image != NULL)
			cst = from_dev(akod_d2h(ctx->dev, image, (uint8_t*)d_out + out_stride * i, image_bytes));
		if (cst != AKO_OK)
		{
			if (image != NULL)
				b->cb.free(image);
			st = cst;
			done = i;
			break;
		}
		b->out[first + i] = image;
	}
	{
		const enum akoStatus sst = from_dev(akod_sync(ctx->dev));
		if (sst != AKO_OK)
		{
			st = sst;
			done = 0;
		}
	}
	if (done != usable && st != AKO_BROKEN_INPUT)
		host_batch_fail(b, first + done, (st != AKO_OK) ? st : AKO_ERROR);
	pool_release(ctx);
	free(blk);
}

struct host_batch_arg
{
	struct host_batch* b;
	int device;
};

static void* host_batch_worker(void* raw)
{
	struct host_batch_arg* a = raw;
	struct host_batch* b = a->b;
	for (;;)
	{
		pthread_mutex_lock(&b->lock);
		const size_t c = b->next_chunk++;
		pthread_mutex_unlock(&b->lock);
		const size_t first = c * b->chunk;
		if (first >= b->n)
			break;
		const size_t count = (b->n - first < b->chunk) ? b->n - first : b->chunk;
		if (b->decode)
			host_batch_decode_chunk(b, first, count, a->device);
		else
			host_batch_encode_chunk(b, first, count, a->device);
	}
	return NULL;
}

/* returns the number of leading images that succeeded; results of later images that did succeed stay valid
 * (out[i] != NULL), failed ones are NULL */
static size_t host_batch_run(struct host_batch* b, enum akoStatus* out_status)
{
	pthread_t helpers[HOST_BATCH_WORKERS];
	struct host_batch_arg args[HOST_BATCH_WORKERS];
	size_t started = 0;
	const size_t chunks = (b->n + b->chunk - 1) / b->chunk;
	pthread_mutex_init(&b->lock, NULL);
	b->next_chunk = 0;
	b->first_failed = b->n;
	b->st = AKO_OK;
	/* worker k works on device k mod n_devices (the caller is worker 0) */
	for (size_t k = 0; k < b->workers; k++)
	{
		args[k].b = b;
		args[k].device = b->devices[k % (size_t)b->n_devices];
	}
	for (size_t k = 1; k < b->workers && k < chunks; k++)
		if (pthread_create(&helpers[started], NULL, host_batch_worker, &args[k]) == 0)
			started++;
	host_batch_worker(&args[0]);
	for (size_t k = 0; k < started; k++)
		pthread_join(helpers[k], NULL);
	pthread_mutex_destroy(&b->lock);
	if (out_status != NULL)
		*out_status = b->st;
	return b->first_failed;
}

static void host_batch_geometry(struct host_batch* b)
{
	/* two workers per device (one chunk in its copy phase while the other computes), two chunks per worker when
	 * the batch allows it, at most HOST_BATCH_CHUNK images each */
	b->n_devices = env_devices(b->devices);
	b->workers = env_size("AKO_B200_BATCH_WORKERS", 2 * (size_t)b->n_devices, 1, HOST_BATCH_WORKERS);
	const size_t most = env_size("AKO_B200_BATCH_CHUNK", HOST_BATCH_CHUNK, 1, HOST_BATCH_CHUNK);
	size_t chunk = (b->n + 2 * b->workers - 1) / (2 * b->workers);
	chunk = (chunk > most) ? most : chunk;
	b->chunk = (chunk < 1) ? 1 : chunk;
}

AKO_API size_t akoB200EncodeBatch(const struct akoCallbacks* c, const struct akoSettings* s, size_t channels, size_t w,
                                  size_t h, size_t n_images, const void* const* in, void** out, size_t* out_sizes,
                                  enum akoStatus* out_status)
{
	struct host_batch b;
	enum akoStatus st = AKO_OK;
	memset(&b, 0, sizeof(b));
	b.cb = (c != NULL) ? *c : akoDefaultCallbacks();
	b.s = (s != NULL) ? *s : akoDefaultSettings();

	if (b.cb.malloc == NULL || b.cb.realloc == NULL || b.cb.free == NULL)
		st = AKO_INVALID_CALLBACKS;
	else if (n_images != 0 && (in == NULL || out == NULL || out_sizes == NULL))
		st = AKO_INVALID_INPUT;
	else
	{
		struct akoSettings v = b.s;
		uint8_t head[16];
		resolve_color(&v);
		st = head_write(channels, w, h, &v, head);
		if (st == AKO_OK && channels == 0)
			st = AKO_INVALID_CHANNELS_NO;
	}
	if (st != AKO_OK || n_images == 0)
	{
		if (out_status != NULL)
			*out_status = st;
		return 0;
	}
	for (size_t i = 0; i < n_images; i++)
	{
		out[i] = NULL;
		out_sizes[i] = 0;
	}
	b.channels = channels;
	b.w = w;
	b.h = h;
	b.n = n_images;
	host_batch_geometry(&b);
	b.in = in;
	b.out = out;
	b.out_sizes = out_sizes;
	return host_batch_run(&b, out_status);
}

AKO_API size_t akoB200DecodeBatch(const struct akoCallbacks* c, size_t n_images, const void* const* in,
                                  const size_t* in_sizes, uint8_t** out, struct akoSettings* out_s, size_t* out_channels,
                                  size_t* out_w, size_t* out_h, enum akoStatus* out_status)
{
	struct host_batch b;
	enum akoStatus st = AKO_OK;
	memset(&b, 0, sizeof(b));
	b.decode = 1;
	b.cb = (c != NULL) ? *c : akoDefaultCallbacks();

	if (b.cb.malloc == NULL || b.cb.realloc == NULL || b.cb.free == NULL)
		st = AKO_INVALID_CALLBACKS;
	else if (n_images != 0 && (in == NULL || in_sizes == NULL || out == NULL || in[0] == NULL))
		st = AKO_INVALID_INPUT;
	else if (n_images != 0 && in_sizes[0] < 16)
		st = AKO_BROKEN_INPUT;
	else if (n_images != 0)
	{
		memset(&b.s, 0, sizeof(b.s));
		st = head_read(in[0], &b.channels, &b.w, &b.h, &b.s);
		if (st == AKO_OK && b.s.wavelet == AKO_WAVELET_NONE && b.s.compression != AKO_COMPRESSION_NONE)
			st = AKO_ERROR; /* see encode_core */
	}
	if (st != AKO_OK || n_images == 0)
	{
		if (out_status != NULL)
			*out_status = st;
		return 0;
	}
	for (size_t i = 0; i < n_images; i++)
		out[i] = NULL;
	b.n = n_images;
	host_batch_geometry(&b);
	b.in = in;
	b.in_sizes = in_sizes;
	b.out = (void**)out;
	if (out_s != NULL)
		*out_s = b.s;
	if (out_channels != NULL)
		*out_channels = b.channels;
	if (out_w != NULL)
		*out_w = b.w;
	if (out_h != NULL)
		*out_h = b.h;
	return host_batch_run(&b, out_status);
}
