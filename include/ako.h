/*
 * ako.h -- public C API of ako-b200, a CUDA (sm_100a) implementation of the Ako
 * wavelet image codec's encode/decode path.
 *
 * This header is ABI-compatible with the reference's library/ako.h (libako 0.2.0,
 * format 2): same exported symbols, same struct layouts (LP64: akoSettings 40 B,
 * akoCallbacks 40 B, akoHead 16 B), same enum values -- the enum values are also
 * wire-visible through the .ako header flags (reference library/head.c:97-102).
 * A program compiled against the reference header links and runs against
 * libako_b200.so unchanged. Device-resident and batched entry points, which the
 * reference does not have, live in ako_b200.h.
 *
 * Each declaration cites the reference interface it replaces (path:line under
 * the reference's library/ directory).
 */
#ifndef AKO_H
#define AKO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ako.h:8-18 */
#define AKO_VERSION_MAJOR 0
#define AKO_VERSION_MINOR 2
#define AKO_VERSION_PATCH 0
#define AKO_FORMAT_VERSION 2

#define AKO_MAX_CHANNELS 16
#define AKO_MAX_WIDTH 4294967295
#define AKO_MAX_HEIGHT 4294967295
#define AKO_MIN_TILES_DIMENSION 8
#define AKO_MAX_TILES_DIMENSION 2147483648

/* ako.h:21-41 -- values returned through the out_status parameters */
enum akoStatus
{
	AKO_OK = 0,
	AKO_ERROR = 1, /* also: any CUDA runtime failure that is not an allocation failure */
	AKO_INVALID_CHANNELS_NO = 2,
	AKO_INVALID_DIMENSIONS = 3,
	AKO_INVALID_TILES_DIMENSIONS = 4,
	AKO_INVALID_WRAP_MODE = 5,
	AKO_INVALID_WAVELET_TRANSFORMATION = 6,
	AKO_INVALID_COLOR_TRANSFORMATION = 7,
	AKO_INVALID_COMPRESSION_METHOD = 8,
	AKO_INVALID_INPUT = 9,
	AKO_INVALID_CALLBACKS = 10,
	AKO_INVALID_MAGIC = 11,
	AKO_UNSUPPORTED_VERSION = 12,
	AKO_NO_ENOUGH_MEMORY = 13, /* also: cudaMalloc failure */
	AKO_INVALID_FLAGS = 14,
	AKO_BROKEN_INPUT = 15
};

/* ako.h:43-49 -- header flag bits 6-7 */
enum akoWavelet
{
	AKO_WAVELET_DD137 = 0,
	AKO_WAVELET_CDF53 = 1,
	AKO_WAVELET_HAAR = 2,
	AKO_WAVELET_NONE = 3
};

/* ako.h:51-58 -- header flag bits 8-9; YCOCG_Q is chosen by the encoder itself when
 * quantization > 0 or gate > 0 (encode.c:59-64) */
enum akoColor
{
	AKO_COLOR_YCOCG = 0,
	AKO_COLOR_SUBTRACT_G = 1,
	AKO_COLOR_NONE = 2,
	AKO_COLOR_YCOCG_Q = 3
};

/* ako.h:60-66 -- header flag bits 4-5 */
enum akoWrap
{
	AKO_WRAP_CLAMP = 0,
	AKO_WRAP_MIRROR = 1,
	AKO_WRAP_REPEAT = 2,
	AKO_WRAP_ZERO = 3
};

/* ako.h:68-73 -- header flag bits 10-11. As in the reference (compression.c:39, :61) the
 * MANBAVARAN value only changes the header flag; the payload is Kagari either way. */
enum akoCompression
{
	AKO_COMPRESSION_KAGARI = 0,
	AKO_COMPRESSION_MANBAVARAN = 1,
	AKO_COMPRESSION_NONE = 2
};

/* ako.h:75-84 -- stage notifications, fired per tile in the reference's order:
 * encode FORMAT, WAVELET, COMPRESSION; decode COMPRESSION, WAVELET, FORMAT */
enum akoEvent
{
	AKO_EVENT_NONE = 0,
	AKO_EVENT_FORMAT_START,
	AKO_EVENT_FORMAT_END,
	AKO_EVENT_WAVELET_START,
	AKO_EVENT_WAVELET_END,
	AKO_EVENT_COMPRESSION_START,
	AKO_EVENT_COMPRESSION_END
};

/* ako.h:86-99 */
struct akoSettings
{
	enum akoWavelet wavelet;
	enum akoColor color;
	enum akoWrap wrap;
	enum akoCompression compression;
	size_t tiles_dimension; /* 0 = one tile, else a power of two >= 8 */

	int quantization; /* <= 0 : lossless step (q = 1 everywhere) */
	int gate;         /* <= 0 : no noise gate */

	int chroma_loss;         /* multiplier-1 applied to q and gate of every channel but the first */
	int discard_non_visible; /* zero colour where alpha == 0 (2 and 4 channel images only) */
};

/* ako.h:101-109 */
struct akoCallbacks
{
	void* (*malloc)(size_t);
	void* (*realloc)(void*, size_t);
	void (*free)(void*);

	void (*events)(size_t tile_no, size_t total_tiles, enum akoEvent, void* events_data);
	void* events_data;
};

/* ako.h:111-127 -- first 16 bytes of every .ako file, little endian */
struct akoHead
{
	uint8_t magic[3]; /* "Ako" */
	uint8_t version;  /* AKO_FORMAT_VERSION */
	uint32_t width;
	uint32_t height;
	uint32_t flags; /* (channels-1) | wrap<<4 | wavelet<<6 | color<<8 | compression<<10 | (log2(tiles)-2)<<12 */
};

/*
 * Encodes an interleaved 8-bit image held in HOST memory into a malloc'ed (callbacks->malloc)
 * .ako blob. Returns the blob size, or 0 with *out_status set. Replaces encode.c:38-232.
 * The work runs on the CUDA device selected by the AKO_CUDA_DEVICE environment variable
 * (default 0); there is no CPU fallback -- without a usable device the call fails with AKO_ERROR.
 */
size_t akoEncodeExt(const struct akoCallbacks*, const struct akoSettings*, size_t channels, size_t image_w,
                    size_t image_h, const void* in, void** out, enum akoStatus* out_status);

/*
 * Decodes a .ako blob held in HOST memory into a malloc'ed interleaved 8-bit image.
 * Returns NULL with *out_status set on failure. Replaces decode.c:38-250.
 */
uint8_t* akoDecodeExt(const struct akoCallbacks*, size_t input_size, const void* in, struct akoSettings* out_s,
                      size_t* out_channels, size_t* out_w, size_t* out_h, enum akoStatus* out_status);

struct akoSettings akoDefaultSettings(void);   /* misc.c:30-47 */
struct akoCallbacks akoDefaultCallbacks(void); /* misc.c:50-62 */
void akoDefaultFree(void*);                    /* misc.c:64-67 */

const char* akoStatusString(enum akoStatus); /* misc.c:71-95 */

int akoVersionMajor(void);  /* version.c:30 */
int akoVersionMinor(void);  /* version.c:36 */
int akoVersionPatch(void);  /* version.c:42 */
int akoFormatVersion(void); /* version.c:48 */

#ifdef __cplusplus
}
#endif
#endif
