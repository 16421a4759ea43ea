/*
 * ako_device.h -- internal thin C-ABI between the host C code (ako_host.c, ako_ext.c) and the
 * CUDA translation unit (ako_device.cu). Plain pointers and sizes only.
 *
 * Division of labour: everything that is geometry, header bits or FLOAT
 * (the quantiser schedule, quantization.c:43-98) is computed on the host and handed to the
 * device as integers inside an akodPlan. The device does integer work only.
 */
#ifndef AKO_DEVICE_H
#define AKO_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AKOD_OK 0
#define AKOD_ERROR 1
#define AKOD_NOMEM 2

#define AKOD_MAX_LEVELS 34
#define AKOD_MAX_CHANNELS 16

/* effective 1-D wavelet of a level (lifting.c:49, :58, :67) */
#define AKOD_DD137 0
#define AKOD_CDF53 1
#define AKOD_HAAR 2

typedef struct akodContext akodContext;

/* One lift level: input cw x ch  ->  four subbands tw x th. Offsets are in int16 units into the
 * tile's coefficient stream. The lift head (q) sits at off_c - 1; B = off_c + tw*th; D = off_c + 2*tw*th. */
typedef struct
{
	uint32_t cw, ch, tw, th;
	int32_t wavelet;
	int16_t q[AKOD_MAX_CHANNELS];
	int16_t g[AKOD_MAX_CHANNELS];
	uint64_t off_c[AKOD_MAX_CHANNELS];
} akodLevel;

/* Everything the device needs to know about one tile shape + settings. level[0] is the finest. */
typedef struct
{
	uint32_t w, h, channels, levels;
	int32_t wrap;
	int32_t wavelet; /* the settings' wavelet (levels carry the effective one) */
	uint32_t lp_w, lp_h;
	uint64_t off_lp[AKOD_MAX_CHANNELS];
	uint64_t stream_len; /* int16 count == akoTileDataSize()*channels/2 */
	akodLevel level[AKOD_MAX_LEVELS];
} akodPlan;

/* Same-shape batches: image i of a buffer lives at base + i*stride (strides in ELEMENTS of the
 * buffer's type). n == 1 and strides 0 for single images. */
typedef struct
{
	uint32_t n;
	uint64_t in_stride;      /* u8 elements between input images / output images */
	uint64_t planes_stride;  /* int16 */
	uint64_t scratch_stride; /* int16 */
	uint64_t stream_stride;  /* int16 */
	uint32_t planes_pitch;   /* int16 elements between the rows of a full-size plane (0: dense, = the width); the cores
	                          * pad it to a multiple of 8 so that every row starts on a 16-byte boundary (TMA row copies) */
	/* Tiles as batch members (encode.c:115-205 / decode.c:113-230 make every tile an independent block): with
	 * n_real != 0 the n members are tiles of n_real images, member v = kk * n_real + i being tile tile_first + kk of
	 * its shape group in image i. Only the interleaved u8 image is addressed through this: tile k of the group
	 * starts tile_y0 + (k / tile_cols) * tile_step rows and tile_x0 + (k % tile_cols) * tile_step pixels into
	 * image i. n_real == 0: a plain batch, member v is image v. */
	uint32_t n_real, tile_cols, tile_first, tile_step, tile_x0, tile_y0;
} akodBatch;

/* The tile grid of an image and where each tile's compressed block and bit count live. Tiles come in up to four
 * shape groups: 0 = full tiles, 1 = the right edge column, 2 = the bottom edge row, 3 = the corner. Tile k of group g
 * of image i has its bit count at bits_base[g] + k * n_images + i and its block at
 * group_base[g] + (k * n_images + i) * group_stride[g] (the order a group's members have in a batch). */
typedef struct
{
	uint32_t tiles_x, tiles_y; /* tiles per row / column of the image */
	uint32_t full_x, full_y;   /* of which full-size */
	uint32_t n_images;
	uint64_t group_base[4];
	uint64_t group_stride[4];
	uint64_t group_cap[4]; /* bytes a block of the group may hold */
	uint64_t bits_base[4];
} akodTiles;

/* raster tile index -> (group, index inside the group) */
static inline void akod_tile_locate(const akodTiles* T, uint32_t t, uint32_t* g, uint32_t* k)
{
	const uint32_t ty = t / T->tiles_x, tx = t - ty * T->tiles_x;
	const uint32_t ex = tx >= T->full_x, ey = ty >= T->full_y;
	*g = ex + 2 * ey;
	*k = (ex && ey) ? 0 : ex ? ty : ey ? tx : ty * T->full_x + tx;
}

/* ---- context ---- */
int akod_context_create(int device, akodContext** out);
void akod_context_destroy(akodContext*);
int akod_device_index(akodContext*);
void* akod_stream(akodContext*);
int akod_sync(akodContext*);
/* blocking != 0: akod_sync sleeps on an event (cudaEventBlockingSync) instead of spinning in the driver. For callers
 * whose waits are milliseconds of PCIe copies and who are many per core (the host-pointer API); a context that
 * serves device-resident calls keeps the spin, whose wake-up is tens of microseconds faster. */
void akod_set_blocking_sync(akodContext*, int blocking);

void* akod_alloc(akodContext*, size_t bytes);
void akod_free(akodContext*, void* d_ptr);
void* akod_pinned_alloc(size_t bytes);
void akod_pinned_free(void* p);
int akod_h2d(akodContext*, void* d_dst, const void* src, size_t bytes);
int akod_d2h(akodContext*, void* dst, const void* d_src, size_t bytes);
int akod_d2d(akodContext*, void* d_dst, const void* d_src, size_t bytes);
int akod_memset(akodContext*, void* d_dst, int value, size_t bytes);
/* count blocks of 'bytes' bytes each: block j from d_src + j*src_stride to d_dst + j*dst_stride (strides in bytes) */
int akod_copy_strided(akodContext*, void* d_dst, uint64_t dst_stride, const void* d_src, uint64_t src_stride, uint64_t bytes,
                      uint64_t count);
/* count blocks of 'bytes' bytes each: block j from d_src + d_off[j] (DEVICE array) to d_dst + j*dst_stride */
int akod_gather(akodContext*, void* d_dst, uint64_t dst_stride, const void* d_src, const uint64_t* d_off, uint64_t bytes,
                uint64_t count);
int akod_fill_words(akodContext*, uint64_t* d_dst, uint64_t value, size_t count);
/* dst / src: device memory or the context's pinned mailbox (directly addressable by kernels); no copy engine */
int akod_copy_words(akodContext*, uint64_t* dst, const uint64_t* src, size_t count);

/* grow-only device workspace slots owned by the context */
enum
{
	AKOD_WS_INPUT = 0, /* staged u8 image / staged blob */
	AKOD_WS_PLANES,
	AKOD_WS_SCRATCH,
	AKOD_WS_STREAM,
	AKOD_WS_KAGARI, /* per-block scan state, tokens */
	AKOD_WS_KAGARI2,
	AKOD_WS_BLOCKS, /* per-tile compressed blocks before assembly */
	AKOD_WS_OUTPUT, /* assembled blob / decoded image */
	AKOD_WS_SMALL,  /* sizes, flags */
	AKOD_WS_SEARCH, /* unquantised coefficient stream kept across the probes of a ratio search */
	AKOD_WS_COUNT
};
int akod_workspace(akodContext*, int slot, size_t bytes, void** out);
/* small page-locked host mailbox owned by the context (>= 64 KiB) */
void* akod_mailbox(akodContext*);

/* ---- profiling ---- */
void akod_profile_enable(akodContext*, int enable);
void akod_profile_reset(akodContext*);
size_t akod_profile_get(akodContext*, size_t cap, const char** names, uint64_t* launches, double* total_ms);
/* algorithmic bytes per kernel name, same order as akod_profile_get */
size_t akod_profile_get_bytes(akodContext*, size_t cap, uint64_t* bytes);
uint64_t akod_launch_count(akodContext*);

/* ---- stages (all asynchronous on the context's stream) ---- */

/* u8 interleaved (row stride in_stride_px pixels) -> dense int16 planes + colour transform */
int akod_format_forward(akodContext*, int discard, int color, uint32_t channels, uint32_t w, uint32_t h,
                        uint64_t in_stride_px, const uint8_t* d_in, int16_t* d_planes, const akodBatch*);
/* dense int16 planes -> u8 interleaved, inverse colour + saturation */
int akod_format_inverse(akodContext*, int color, uint32_t channels, uint32_t w, uint32_t h, uint64_t out_stride_px,
                        const int16_t* d_planes, uint8_t* d_out, const akodBatch*);

/* full pyramid; d_planes (channels*w*h) is destroyed; d_scratch needs channels*ceil(w/2)*ceil(h/2) */
int akod_lift(akodContext*, const akodPlan*, int16_t* d_planes, int16_t* d_scratch, int16_t* d_stream,
              const akodBatch*);
/* akod_format_forward followed by akod_lift; the two are one kernel on level 0 when the image allows (lift_strip4.cuh):
 * akod_format_lift_fuses() tells (1 / 0) */
int akod_format_lift_fuses(akodContext*, uint32_t channels, uint32_t w, uint32_t h, uint64_t in_stride_px,
                           const uint8_t* d_in, const akodPlan*, int16_t* d_planes, int16_t* d_scratch, int16_t* d_stream,
                           const akodBatch*);
int akod_format_lift(akodContext*, int discard, int color, uint32_t channels, uint32_t w, uint32_t h,
                     uint64_t in_stride_px, const uint8_t* d_in, const akodPlan*, int16_t* d_planes, int16_t* d_scratch,
                     int16_t* d_stream, const akodBatch*);
/* full inverse pyramid into d_planes; d_scratch as above; d_stream untouched */
int akod_unlift(akodContext*, const akodPlan*, const int16_t* d_stream, int16_t* d_planes, int16_t* d_scratch,
                const akodBatch*);

/* Kagari: n_values int16 per image -> bytes at d_out (+ i*out_stride bytes per image), 4-byte aligned.
 * d_bits[i] (uint64, device) receives the exact bit length of image i's stream. Nothing is written for an
 * image whose stream would not fit in out_cap bytes (out_cap a multiple of 4); the caller sees that in d_bits. */
int akod_kagari_encode(akodContext*, uint64_t n_values, const int16_t* d_in, uint64_t in_stride, uint8_t* d_out,
                       uint64_t out_stride, uint64_t out_cap, uint64_t* d_bits, uint32_t n_images);

/* Size probe: d_bits[i] = exact bit length of image i's Kagari stream, nothing is packed. */
int akod_kagari_bits(akodContext*, uint64_t n_values, const int16_t* d_in, uint64_t in_stride, uint64_t* d_bits,
                     uint32_t n_images);
/* d_in: a stream lifted with q = 1, gate = 0 on every level. d_out: the stream akod_lift produces with 'plan'. */
int akod_requantize(akodContext*, const akodPlan*, const int16_t* d_in, int16_t* d_out);

/* Kagari decode of n_images blocks. Block i: d_size[i] bytes at d_in + d_off[i] (DEVICE arrays).
 * d_result[i] (uint64, device): bytes consumed, or 0 when the block is malformed. */
int akod_kagari_decode(akodContext*, uint64_t n_values, const uint8_t* d_in, const uint64_t* d_off,
                       const uint64_t* d_size, uint64_t max_in_size /* host copy of max(d_size) */, int16_t* d_out,
                       uint64_t out_stride, uint64_t* d_result, uint32_t n_images);

/* Container assembly on the device (encode.c:170-182 without the CPU pass), for n_images same-shape images:
 * out_i = head16 | for each tile t in raster order: [u32 size][bytes of tile t]. Sizes come from d_bits, blocks
 * from d_blocks, both laid out as akodTiles says; an image with a tile over its capacity gets d_total[i] = 0 and
 * nothing written. d_total[i] = blob size. */
int akod_assemble(akodContext*, const uint8_t head16[16], const akodTiles* tiles, const uint8_t* d_blocks,
                  const uint64_t* d_bits, int with_heads, uint8_t* d_out, uint64_t out_stride, uint64_t* d_total);
/* walk the block heads of a device-resident blob: d_off[t] = byte offset of tile t's payload,
 * d_size[t] = its block_size; stops (size 0) if it would run past input_size */
int akod_walk_blocks(akodContext*, const uint8_t* d_blob, uint64_t input_size, uint32_t n_tiles, uint64_t* d_off,
                     uint64_t* d_size);
/* the same for n blobs at d_blobs + i*stride with d_input_size[i] bytes; d_off / d_size are [n][n_tiles] */
int akod_walk_blocks_batch(akodContext*, const uint8_t* d_blobs, uint64_t stride, const uint64_t* d_input_size,
                           uint32_t n_tiles, uint32_t n, uint64_t* d_off, uint64_t* d_size);

#ifdef __cplusplus
}
#endif
#endif
