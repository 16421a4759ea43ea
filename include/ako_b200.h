/*
 * ako_b200.h -- additive extension of the Ako C API for device-resident and batched use.
 *
 * The reference (library/ako.h) only has host-pointer, one-image entry points. Everything
 * here is NEW surface; per image it produces exactly what akoEncodeExt / akoDecodeExt
 * produce (same .ako bytes, same pixels). Pointers prefixed d_ are CUDA device pointers on
 * the context's device; all other pointers are host pointers. Plain C types only.
 *
 * Threading: a context serialises its own work on one CUDA stream. Use one context per
 * host thread; different contexts (on the same or different GPUs) run concurrently.
 */
#ifndef AKO_B200_H
#define AKO_B200_H

#include "ako.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct akoB200Context akoB200Context;

/* ---- contexts, memory, timing ------------------------------------------------------ */

/* device < 0 : use $AKO_CUDA_DEVICE (default 0). NULL on failure (*out_status says why). */
akoB200Context* akoB200ContextCreate(int device, enum akoStatus* out_status);
void akoB200ContextDestroy(akoB200Context*);

/* The cudaStream_t all work of this context is issued on (cast to void*), so callers can
 * bracket calls with their own CUDA events. */
void* akoB200ContextStream(akoB200Context*);
enum akoStatus akoB200Synchronize(akoB200Context*);

/* Thin wrappers so plain-C callers need no CUDA headers. */
void* akoB200DeviceAlloc(akoB200Context*, size_t bytes);
void akoB200DeviceFree(akoB200Context*, void* d_ptr);
void* akoB200PinnedAlloc(size_t bytes); /* page-locked host memory */
void akoB200PinnedFree(void* ptr);
enum akoStatus akoB200CopyToDevice(akoB200Context*, void* d_dst, const void* src, size_t bytes);   /* async */
enum akoStatus akoB200CopyToHost(akoB200Context*, void* dst, const void* d_src, size_t bytes);     /* async */

/* akoCallbacks whose malloc/realloc/free hand out page-locked host memory, so that the
 * blob / image returned by akoEncodeExt / akoDecodeExt arrives by DMA at full PCIe rate.
 * Free what they return with callbacks.free (== akoB200PinnedFree). */
struct akoCallbacks akoB200PinnedCallbacks(void);

/* Per-kernel accounting. When enabled every kernel launch is bracketed by CUDA events on the
 * context's stream; read the totals back with akoB200ProfileGet. Counting launches is always on. */
void akoB200ProfileEnable(akoB200Context*, int enable);
void akoB200ProfileReset(akoB200Context*);
/* number of distinct kernels seen; fills up to cap entries. names[i] points at static strings. */
size_t akoB200ProfileGet(akoB200Context*, size_t cap, const char** names, uint64_t* launches, double* total_ms);
/* Algorithmic bytes (compulsory reads + writes) the launches of each kernel name moved, in the
 * order of akoB200ProfileGet; accounted whether or not timing is enabled. */
size_t akoB200ProfileGetBytes(akoB200Context*, size_t cap, uint64_t* bytes);
uint64_t akoB200LaunchCount(akoB200Context*);

/* ---- whole codec, device resident ---------------------------------------------------- */

/* Upper bound of the .ako size for this image: what d_out must be able to hold. */
size_t akoB200EncodeBound(const struct akoSettings*, size_t channels, size_t image_w, size_t image_h);

/* d_in : interleaved u8 image on the device. d_out : receives the complete .ako blob (header
 * included). Returns the blob size (0 on failure). Same checks, same order, same statuses as
 * akoEncodeExt (encode.c:38-232). */
size_t akoB200EncodeDevice(akoB200Context*, const struct akoSettings*, size_t channels, size_t image_w,
                           size_t image_h, const void* d_in, void* d_out, size_t out_capacity,
                           enum akoStatus* out_status);

/* d_in : .ako blob on the device (input_size bytes). head16 : host copy of its first 16 bytes, or
 * NULL to let the call fetch them. d_out : receives w*h*channels interleaved u8 (capacity checked).
 * Same statuses as akoDecodeExt (decode.c:38-250). */
enum akoStatus akoB200DecodeDevice(akoB200Context*, size_t input_size, const void* d_in, const void* head16,
                                   void* d_out, size_t out_capacity, struct akoSettings* out_s,
                                   size_t* out_channels, size_t* out_w, size_t* out_h);

/* Batches of same-shape images. Image i is at d_in + i*in_stride (in_stride >= w*h*channels);
 * blob i is written at d_out + i*out_stride and its size stored in out_sizes[i] (host array).
 * Returns the number of images encoded before the first failure (== n_images on success). */
size_t akoB200EncodeBatchDevice(akoB200Context*, const struct akoSettings*, size_t channels, size_t image_w,
                                size_t image_h, size_t n_images, const void* d_in, size_t in_stride, void* d_out,
                                size_t out_stride, size_t* out_sizes, enum akoStatus* out_status);

/* Blob i is at d_in + i*in_stride with in_sizes[i] bytes; all blobs must describe the same
 * channels/w/h/settings (checked against blob 0). Image i is written at d_out + i*out_stride. */
size_t akoB200DecodeBatchDevice(akoB200Context*, size_t n_images, const void* d_in, size_t in_stride,
                                const size_t* in_sizes, void* d_out, size_t out_stride,
                                enum akoStatus* out_status);

/* ---- batches of HOST images / HOST blobs ---------------------------------------------------- */

/* n_images same-shape images at host pointers in[i] -> out[i] = blob i (allocated with callbacks->malloc, the
 * caller frees each), out_sizes[i] its size; per image exactly what akoEncodeExt returns. Inside, the batch runs in
 * chunks of a few images through the batched kernels on several streams, so that the upload of one chunk overlaps
 * the kernels of another and the read-back of a third. Pinned input buffers (akoB200PinnedAlloc) and
 * akoB200PinnedCallbacks() make every copy a DMA. Returns the number of leading images that succeeded
 * (== n_images when all did; *out_status then AKO_OK, else the status of the first failure). Images after a
 * failure that did succeed keep their blobs; failed ones have out[i] == NULL. Events are not fired. */
size_t akoB200EncodeBatch(const struct akoCallbacks*, const struct akoSettings*, size_t channels, size_t image_w,
                          size_t image_h, size_t n_images, const void* const* in, void** out, size_t* out_sizes,
                          enum akoStatus* out_status);

/* n_images blobs (host pointers in[i], in_sizes[i] bytes) that all describe the channels / dimensions / settings of
 * blob 0 -> out[i] = w*h*channels interleaved u8 (callbacks->malloc); per image what akoDecodeExt returns. A blob
 * of another shape fails with AKO_INVALID_INPUT. Same return convention as akoB200EncodeBatch. */
size_t akoB200DecodeBatch(const struct akoCallbacks*, size_t n_images, const void* const* in, const size_t* in_sizes,
                          uint8_t** out, struct akoSettings* out_s, size_t* out_channels, size_t* out_w, size_t* out_h,
                          enum akoStatus* out_status);

/* ---- ratio search ------------------------------------------------------------------------ */

/* What the encoder tool's EncodePass does (tools/akoenc.cpp:111-213): search the quantisation that brings the
 * .ako size within 4 % of image_bytes / ratio and return THAT blob -- same passes, same decisions, same bytes as
 * the tool calling akoEncodeExt up to a dozen times. Here the image is uploaded once, the colour transform and the
 * wavelet run once per colour model, and every pass only re-quantises the device-resident coefficient stream and
 * measures its Kagari length; only the winning pass is packed. ratio 0: plain akoEncodeExt; ratio 1: lossless.
 * Host pointers, ownership and statuses as akoEncodeExt. *out_quantization: the quantization of the returned blob;
 * *out_passes: the number of akoEncodeExt calls the tool would have made. */
size_t akoB200EncodeRatio(const struct akoCallbacks*, const struct akoSettings*, int ratio, size_t channels,
                          size_t image_w, size_t image_h, const void* in, void** out, int* out_quantization,
                          size_t* out_passes, enum akoStatus* out_status);

/* ---- single stages, device resident (parity tests, DWT-only sweeps) -------------------- */

/* Bytes of the coefficient stream of one tile/image: akoTileDataSize()*channels, misc.c:117-149 */
size_t akoB200StreamSize(size_t channels, size_t tile_w, size_t tile_h);

/* format.c:64-135. d_planes: channels dense int16 planes of w*h. */
enum akoStatus akoB200FormatForward(akoB200Context*, const struct akoSettings*, size_t channels, size_t w, size_t h,
                                    size_t in_stride_px, const void* d_in, int16_t* d_planes);
/* format.c:244-311. Does not modify d_planes. */
enum akoStatus akoB200FormatInverse(akoB200Context*, enum akoColor, size_t channels, size_t w, size_t h,
                                    size_t out_stride_px, const int16_t* d_planes, void* d_out);

/* lifting.c:171-292 (akoLift): dense planes -> coefficient stream, quantised and gated.
 * d_planes is used as scratch and destroyed. */
enum akoStatus akoB200Lift(akoB200Context*, const struct akoSettings*, size_t channels, size_t w, size_t h,
                           int16_t* d_planes, int16_t* d_stream);
/* lifting.c:295-304 (akoUnlift): coefficient stream -> dense planes. d_stream is not modified. */
enum akoStatus akoB200Unlift(akoB200Context*, const struct akoSettings*, size_t channels, size_t w, size_t h,
                             const int16_t* d_stream, int16_t* d_planes);

/* kagari.c:228-298. Returns bytes written to d_out, 0 if they would not be < out_capacity
 * (the reference's rule, kagari.c:65-68, :93-107) or on error. */
size_t akoB200KagariEncode(akoB200Context*, size_t n_values, const int16_t* d_in, void* d_out, size_t out_capacity,
                           enum akoStatus* out_status);
/* kagari.c:301-366. Returns bytes consumed (== in_size for a well-formed block), 0 on failure. */
size_t akoB200KagariDecode(akoB200Context*, size_t n_values, size_t in_size, const void* d_in, int16_t* d_out,
                           enum akoStatus* out_status);

#ifdef __cplusplus
}
#endif
#endif
