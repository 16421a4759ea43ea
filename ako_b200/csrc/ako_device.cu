// ako_device.cu -- the CUDA translation unit of libako_b200: context management, kernel launchers
// and the container kernels, exported through the internal C-ABI of ako_device.h.
// Compiled for sm_100a only (no other architecture is built; see Makefile).

#include <stdlib.h>

#include "common.cuh"
#include "format.cuh"
#include "kagari_dec.cuh"
#include "kagari_enc.cuh"
#include "lift.cuh"
#include "lift_small.cuh"
#include "lift_strip.cuh"
#include "lift_strip4.cuh"
#include "unlift_strip.cuh"

// ------------------------------------------------------------------------------------------------
// context

// CUDA's current device is per host thread, a context's stream and buffers live on ONE device: every entry point
// that launches, copies or synchronises first makes the context's device current on the calling thread (contexts are
// pooled by the host API and handed to whichever thread calls next).
static inline void akod_use(const akodContext* c)
{
	int cur = -1;
	if (cudaGetDevice(&cur) != cudaSuccess || cur != c->device)
		cudaSetDevice(c->device);
}

extern "C" int akod_context_create(int device, akodContext** out)
{
	*out = nullptr;
	int count = 0;
	AKOD_TRY(cudaGetDeviceCount(&count));
	if (device < 0 || device >= count)
		return AKOD_ERROR;
	AKOD_TRY(cudaSetDevice(device));

	akodContext* c = new akodContext();
	c->device = device;
	c->profiling = false;
	c->launch_count = 0;
	c->next_bytes = 0;
	c->small_attr_done = false;
	c->strip4_attr_done = 0;
	c->mailbox = nullptr;
	c->sync_event = nullptr;
	c->blocking_sync = false;
	for (int i = 0; i < AKOD_WS_COUNT; i++)
	{
		c->ws[i] = nullptr;
		c->ws_size[i] = 0;
	}
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
	{
		delete c;
		return AKOD_ERROR;
	}
	c->sm_count = prop.multiProcessorCount;
	if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaEventCreateWithFlags(&c->sync_event, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess ||
	    cudaHostAlloc(&c->mailbox, 1 << 16, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
	{
		delete c;
		return AKOD_ERROR;
	}
	*out = c;
	return AKOD_OK;
}

static void akod_collect_pending(akodContext* c)
{
	if (c->pending.empty())
		return;
	cudaStreamSynchronize(c->stream);
	for (akodPending& p : c->pending)
	{
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess)
			c->prof[p.entry].ms += ms;
		c->event_pool.push_back(p.a);
		c->event_pool.push_back(p.b);
	}
	c->pending.clear();
}

extern "C" void akod_context_destroy(akodContext* c)
{
	if (!c)
		return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	akod_collect_pending(c);
	for (cudaEvent_t e : c->event_pool)
		cudaEventDestroy(e);
	for (int i = 0; i < AKOD_WS_COUNT; i++)
		if (c->ws[i])
			cudaFree(c->ws[i]);
	if (c->mailbox)
		cudaFreeHost(c->mailbox);
	if (c->sync_event)
		cudaEventDestroy(c->sync_event);
	cudaStreamDestroy(c->stream);
	delete c;
}

extern "C" int akod_device_index(akodContext* c)
{
	return c->device;
}

extern "C" void* akod_stream(akodContext* c)
{
	return (void*)c->stream;
}

extern "C" int akod_sync(akodContext* c)
{
	akod_use(c);
	if (c->blocking_sync)
	{
		AKOD_TRY(cudaEventRecord(c->sync_event, c->stream));
		AKOD_TRY(cudaEventSynchronize(c->sync_event));
		return AKOD_OK;
	}
	AKOD_TRY(cudaStreamSynchronize(c->stream));
	return AKOD_OK;
}

extern "C" void akod_set_blocking_sync(akodContext* c, int blocking)
{
	c->blocking_sync = blocking != 0;
}

extern "C" void* akod_alloc(akodContext* c, size_t bytes)
{
	void* p = nullptr;
	cudaSetDevice(c->device);
	if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess)
	{
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

extern "C" void akod_free(akodContext* c, void* p)
{
	if (p)
	{
		cudaSetDevice(c->device);
		cudaFree(p);
	}
}

extern "C" void* akod_pinned_alloc(size_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) // usable from every device
	{
		cudaGetLastError();
		return nullptr;
	}
	return p;
}

extern "C" void akod_pinned_free(void* p)
{
	if (p)
		cudaFreeHost(p);
}

// Copies of up to AKOD_KERNEL_COPY_MAX bytes between device memory and PAGE-LOCKED host memory are done by a kernel
// (page-locked memory is addressable by kernels), everything else by the copy engines. A .ako blob is tens of
// kilobytes: as a DMA request it queues behind the 16 MB image copies that other contexts have in flight on the
// same engine, as a kernel it does not. Visible to the host after the stream is synchronised, like the DMA.
constexpr size_t AKOD_KERNEL_COPY_MAX = (size_t)2 << 20;

__global__ void __launch_bounds__(256) k_copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t n)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
	if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0)
	{
		const size_t n16 = n >> 4;
		for (size_t i = t; i < n16; i += stride)
			reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
		for (size_t i = (n16 << 4) + t; i < n; i += stride)
			dst[i] = src[i];
	}
	else
		for (size_t i = t; i < n; i += stride)
			dst[i] = src[i];
}

static bool akod_host_pinned(const void* p)
{
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
	{
		cudaGetLastError();
		return false;
	}
	return a.type == cudaMemoryTypeHost;
}

static int akod_small_copy(akodContext* c, void* d, const void* s, size_t n)
{
	const unsigned blocks = (unsigned)((n / 16 + 255) / 256 < 128 ? (n / 16 + 255) / 256 + 1 : 128);
	AKOD_LAUNCH(c, "copy_bytes", k_copy_bytes, blocks, 256, 0, (uint8_t*)d, (const uint8_t*)s, n);
	return AKOD_OK;
}

extern "C" int akod_h2d(akodContext* c, void* d, const void* s, size_t n)
{
	akod_use(c);
	if (n != 0 && n <= AKOD_KERNEL_COPY_MAX && akod_host_pinned(s))
		return akod_small_copy(c, d, s, n);
	AKOD_TRY(cudaMemcpyAsync(d, s, n, cudaMemcpyDefault, c->stream)); // unified addressing: the driver looks the pointers up
	return AKOD_OK;
}

extern "C" int akod_d2h(akodContext* c, void* d, const void* s, size_t n)
{
	akod_use(c);
	if (n != 0 && n <= AKOD_KERNEL_COPY_MAX && akod_host_pinned(d))
		return akod_small_copy(c, d, s, n);
	AKOD_TRY(cudaMemcpyAsync(d, s, n, cudaMemcpyDefault, c->stream));
	return AKOD_OK;
}

extern "C" int akod_d2d(akodContext* c, void* d, const void* s, size_t n)
{
	akod_use(c);
	AKOD_TRY(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, c->stream));
	return AKOD_OK;
}

extern "C" int akod_memset(akodContext* c, void* d, int v, size_t n)
{
	akod_use(c);
	AKOD_TRY(cudaMemsetAsync(d, v, n, c->stream));
	return AKOD_OK;
}

// block j (blockIdx.y, folded over gridDim.z) of 'bytes' bytes from src + off(j) to dst + j*dst_stride
__global__ void __launch_bounds__(256)
    k_copy_blocks(uint8_t* __restrict__ dst, uint64_t dst_stride, const uint8_t* __restrict__ src, uint64_t src_stride,
                  const uint64_t* __restrict__ src_off, uint64_t bytes, uint64_t count)
{
	const uint64_t j = (uint64_t)blockIdx.z * gridDim.y + blockIdx.y;
	if (j >= count)
		return;
	uint8_t* d = dst + j * dst_stride;
	const uint8_t* s = src + (src_off ? src_off[j] : j * src_stride);
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
	if ((((uintptr_t)d | (uintptr_t)s) & 15) == 0)
	{
		const size_t n16 = bytes >> 4;
		for (size_t i = t; i < n16; i += stride)
			reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
		for (size_t i = (n16 << 4) + t; i < bytes; i += stride)
			d[i] = s[i];
	}
	else
		for (size_t i = t; i < bytes; i += stride)
			d[i] = s[i];
}

static int akod_copy_blocks(akodContext* c, void* d, uint64_t dst_stride, const void* s, uint64_t src_stride,
                            const uint64_t* d_off, uint64_t bytes, uint64_t count)
{
	akod_use(c);
	if (count == 0 || bytes == 0)
		return AKOD_OK;
	const uint64_t per = (bytes / 16 + 255) / 256 + 1;
	const unsigned gx = (unsigned)(per < 32 ? per : 32);
	const unsigned gy = (unsigned)(count < 32768 ? count : 32768), gz = (unsigned)((count + gy - 1) / gy);
	if (gz > 65535)
		return AKOD_ERROR;
	AKOD_LAUNCH(c, "copy_blocks", k_copy_blocks, dim3(gx, gy, gz), 256, 0, (uint8_t*)d, dst_stride, (const uint8_t*)s,
	            src_stride, d_off, bytes, count);
	return AKOD_OK;
}

extern "C" int akod_copy_strided(akodContext* c, void* d, uint64_t dst_stride, const void* s, uint64_t src_stride,
                                 uint64_t bytes, uint64_t count)
{
	return akod_copy_blocks(c, d, dst_stride, s, src_stride, nullptr, bytes, count);
}

extern "C" int akod_gather(akodContext* c, void* d, uint64_t dst_stride, const void* s, const uint64_t* d_off, uint64_t bytes,
                           uint64_t count)
{
	return akod_copy_blocks(c, d, dst_stride, s, 0, d_off, bytes, count);
}

__global__ void k_fill_words(uint64_t* dst, uint64_t value, size_t count)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count)
		dst[i] = value;
}

extern "C" int akod_fill_words(akodContext* c, uint64_t* d, uint64_t value, size_t count)
{
	akod_use(c);
	AKOD_LAUNCH(c, "fill_words", k_fill_words, (unsigned)((count + 255) / 256), 256, 0, d, value, count);
	return AKOD_OK;
}

// Small word arrays between the device and the context's pinned mailbox WITHOUT the copy engines: page-locked host
// memory is directly addressable by kernels (unified addressing), so a tiny kernel moves the words. A DMA request of
// a few bytes queues behind whatever 16 MB image copies other contexts have in flight on the same engine; a kernel
// does not. Device writes to host memory are visible to the host once the stream has been synchronised.
__global__ void k_copy_words(uint64_t* __restrict__ dst, const uint64_t* __restrict__ src, size_t count)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
		dst[i] = src[i];
}

extern "C" int akod_copy_words(akodContext* c, uint64_t* dst, const uint64_t* src, size_t count)
{
	akod_use(c);
	if (count == 0)
		return AKOD_OK;
	const unsigned blocks = (unsigned)((count + 255) / 256 < 64 ? (count + 255) / 256 : 64);
	AKOD_LAUNCH(c, "copy_words", k_copy_words, blocks, 256, 0, dst, src, count);
	return AKOD_OK;
}

extern "C" int akod_workspace(akodContext* c, int slot, size_t bytes, void** out)
{
	akod_use(c);
	*out = nullptr;
	if (slot < 0 || slot >= AKOD_WS_COUNT)
		return AKOD_ERROR;
	if (c->ws_size[slot] < bytes)
	{
		cudaSetDevice(c->device);
		// earlier work may still be using the old buffer
		AKOD_TRY(cudaStreamSynchronize(c->stream));
		if (c->ws[slot])
			cudaFree(c->ws[slot]);
		c->ws[slot] = nullptr;
		c->ws_size[slot] = 0;
		const size_t rounded = (bytes + ((size_t)1 << 20)) & ~(((size_t)1 << 20) - 1);
		AKOD_TRY(cudaMalloc(&c->ws[slot], rounded));
		c->ws_size[slot] = rounded;
	}
	*out = c->ws[slot];
	return AKOD_OK;
}

extern "C" void* akod_mailbox(akodContext* c)
{
	return c->mailbox;
}

extern "C" void akod_profile_enable(akodContext* c, int enable)
{
	akod_collect_pending(c);
	c->profiling = enable != 0;
}

extern "C" void akod_profile_reset(akodContext* c)
{
	akod_use(c);
	akod_collect_pending(c);
	c->prof.clear();
	c->launch_count = 0;
}

extern "C" size_t akod_profile_get(akodContext* c, size_t cap, const char** names, uint64_t* launches, double* ms)
{
	akod_use(c);
	akod_collect_pending(c);
	for (size_t i = 0; i < c->prof.size() && i < cap; i++)
	{
		names[i] = c->prof[i].name;
		launches[i] = c->prof[i].launches;
		ms[i] = c->prof[i].ms;
	}
	return c->prof.size();
}

extern "C" size_t akod_profile_get_bytes(akodContext* c, size_t cap, uint64_t* bytes)
{
	for (size_t i = 0; i < c->prof.size() && i < cap; i++)
		bytes[i] = c->prof[i].bytes;
	return c->prof.size();
}

extern "C" uint64_t akod_launch_count(akodContext* c)
{
	return c->launch_count;
}

// grid for grid-stride streaming kernels: a whole number of waves of the SM count
static inline unsigned akod_stream_grid(akodContext* c, uint64_t items, unsigned block, unsigned ctas_per_sm)
{
	const uint64_t need = (items + block - 1) / block;
	const uint64_t cap = (uint64_t)c->sm_count * ctas_per_sm;
	return (unsigned)(need < cap ? (need ? need : 1) : cap);
}

// ------------------------------------------------------------------------------------------------
// format

static FmtTiles fmt_tiles(const akodBatch* b)
{
	FmtTiles t;
	memset(&t, 0, sizeof(t));
	if (b && b->n_real)
	{
		t.n_real = b->n_real;
		t.cols = b->tile_cols ? b->tile_cols : 1;
		t.first = b->tile_first;
		t.step = b->tile_step;
		t.x0 = b->tile_x0;
		t.y0 = b->tile_y0;
	}
	return t;
}

// the 128-bit path needs every member's first pixel on a 16-byte boundary (4 channels: a multiple of 4 pixels)
static bool fmt_tiles_aligned(const FmtTiles& t)
{
	return t.n_real == 0 || ((t.step % 4) == 0 && (t.x0 % 4) == 0);
}

extern "C" int akod_format_forward(akodContext* c, int discard, int color, uint32_t channels, uint32_t w, uint32_t h,
                                   uint64_t in_stride_px, const uint8_t* d_in, int16_t* d_planes, const akodBatch* b)
{
	akod_use(c);
	const uint32_t n = b ? b->n : 1;
	const uint64_t in_is = b ? b->in_stride : 0, pl_is = b ? b->planes_stride : 0;
	const FmtTiles ft = fmt_tiles(b);
	const uint32_t pitch = (b && b->planes_pitch) ? b->planes_pitch : w;
	const bool fast = channels == 4 && (w % 8) == 0 && (pitch % 8) == 0 && (in_stride_px % 4) == 0 && ((uintptr_t)d_in % 16) == 0 &&
	                  ((uintptr_t)d_planes % 16) == 0 && (in_is % 16) == 0 && (pl_is % 8) == 0 && fmt_tiles_aligned(ft);
	// RGB8: 8 pixels are 24 bytes, 8-byte aligned when every row and member starts on a multiple of 8 pixels
	const bool fast3 = channels == 3 && (w % 8) == 0 && (pitch % 8) == 0 && (in_stride_px % 8) == 0 && ((uintptr_t)d_in % 8) == 0 &&
	                   ((uintptr_t)d_planes % 16) == 0 && (in_is % 8) == 0 && (pl_is % 8) == 0 &&
	                   (ft.n_real == 0 || ((ft.step % 8) == 0 && (ft.x0 % 8) == 0));
	AKOD_BYTES(c, (uint64_t)3 * w * h * channels * n); // u8 in, int16 out
	// the members of a batch ride in gridDim.y; a member of few pixels gets a grid of few CTAs
	const unsigned per_sm = n >= 64 ? 1 : 8;
	if (fast)
	{
		const dim3 grid(akod_stream_grid(c, (uint64_t)(w / 8) * h, 256, per_sm), n);
		AKOD_LAUNCH(c, "format_fwd_rgba8x8", k_format_fwd_rgba8x8, grid, 256, 0, d_in, d_planes, w, h, in_stride_px, color,
		            discard, in_is, pl_is, ft, pitch);
	}
	else if (fast3)
	{
		const dim3 grid(akod_stream_grid(c, (uint64_t)(w / 8) * h, 256, per_sm), n);
		AKOD_LAUNCH(c, "format_fwd_rgb8x8", k_format_fwd_rgb8x8, grid, 256, 0, d_in, d_planes, w, h, in_stride_px, color, in_is,
		            pl_is, ft, pitch);
	}
	else
	{
		const dim3 grid(akod_stream_grid(c, (uint64_t)w * h, 256, per_sm), n);
		AKOD_LAUNCH(c, "format_fwd_generic", k_format_fwd_generic, grid, 256, 0, d_in, d_planes, channels, w, h,
		            in_stride_px, color, discard, in_is, pl_is, ft, pitch);
	}
	return AKOD_OK;
}

extern "C" int akod_format_inverse(akodContext* c, int color, uint32_t channels, uint32_t w, uint32_t h,
                                   uint64_t out_stride_px, const int16_t* d_planes, uint8_t* d_out, const akodBatch* b)
{
	akod_use(c);
	const uint32_t n = b ? b->n : 1;
	const uint64_t out_is = b ? b->in_stride : 0, pl_is = b ? b->planes_stride : 0;
	const FmtTiles ft = fmt_tiles(b);
	const uint32_t pitch = (b && b->planes_pitch) ? b->planes_pitch : w;
	const bool fast = channels == 4 && (w % 8) == 0 && (pitch % 8) == 0 && (out_stride_px % 4) == 0 && ((uintptr_t)d_out % 16) == 0 &&
	                  ((uintptr_t)d_planes % 16) == 0 && (out_is % 16) == 0 && (pl_is % 8) == 0 && fmt_tiles_aligned(ft);
	const bool fast3 = channels == 3 && (w % 8) == 0 && (pitch % 8) == 0 && (out_stride_px % 8) == 0 && ((uintptr_t)d_out % 8) == 0 &&
	                   ((uintptr_t)d_planes % 16) == 0 && (out_is % 8) == 0 && (pl_is % 8) == 0 &&
	                   (ft.n_real == 0 || ((ft.step % 8) == 0 && (ft.x0 % 8) == 0));
	AKOD_BYTES(c, (uint64_t)3 * w * h * channels * n); // int16 in, u8 out
	const unsigned per_sm = n >= 64 ? 1 : 8;
	if (fast)
	{
		const dim3 grid(akod_stream_grid(c, (uint64_t)(w / 8) * h, 256, per_sm), n);
		AKOD_LAUNCH(c, "format_inv_rgba8x8", k_format_inv_rgba8x8, grid, 256, 0, d_planes, d_out, w, h, out_stride_px, color,
		            pl_is, out_is, ft, pitch);
	}
	else if (fast3)
	{
		const dim3 grid(akod_stream_grid(c, (uint64_t)(w / 8) * h, 256, per_sm), n);
		AKOD_LAUNCH(c, "format_inv_rgb8x8", k_format_inv_rgb8x8, grid, 256, 0, d_planes, d_out, w, h, out_stride_px, color, pl_is,
		            out_is, ft, pitch);
	}
	else
	{
		const dim3 grid(akod_stream_grid(c, (uint64_t)w * h, 256, per_sm), n);
		AKOD_LAUNCH(c, "format_inv_generic", k_format_inv_generic, grid, 256, 0, d_planes, d_out, channels, w, h,
		            out_stride_px, color, pl_is, out_is, ft, pitch);
	}
	return AKOD_OK;
}

// ------------------------------------------------------------------------------------------------
// lifting

template <int WL>
static int launch_lift_level(akodContext* c, const LiftParams& p, uint32_t n_images)
{
	const dim3 grid((p.tw + LIFT_TW - 1) / LIFT_TW, (p.th + LIFT_TH - 1) / LIFT_TH, p.channels * n_images);
	static const char* const names[3] = {"lift_dd137", "lift_cdf53", "lift_haar"};
	AKOD_BYTES(c, (uint64_t)4 * p.cw * p.ch * p.channels * n_images);
	AKOD_LAUNCH(c, names[WL], k_lift_level<WL>, grid, LIFT_THREADS, lift_smem_bytes<WL>(), p);
	return AKOD_OK;
}

// A level with a wrap mode other than CLAMP on the strip kernels: the strip kernel does the plane as CLAMP, the general
// kernel then does the FRAME_BAND rows / columns along the plane's edge again with the real wrap mode, as thin tiles
// (lift_tile_origin). Worth it from a few tiles each way (w, h = the level's size in coefficients): against the general
// kernel alone (about five strip kernels' time per coefficient) the strip pass costs w h, the frame about 5 x 8 x 2 (1.6 w +
// 2.2 h) with its halos.
static inline bool frame_worth(uint32_t w, uint32_t h)
{
	return w >= 2 * LIFT_TW && h >= 2 * LIFT_TH;
}

template <int WL>
static int launch_lift_frame(akodContext* c, const LiftParams& p_in, uint32_t n_images)
{
	LiftParams p = p_in;
	static const char* const names[3] = {"lift_frame_dd137", "lift_frame_cdf53", "lift_frame_haar"};
	p.frame = FRAME_ROWS;
	const dim3 grid_r((p.tw + LIFT_TW - 1) / LIFT_TW, 2, p.channels * n_images);
	AKOD_BYTES(c, (uint64_t)4 * 4 * LIFT_TW * FRAME_BAND * grid_r.x * grid_r.y * grid_r.z);
	AKOD_LAUNCH(c, names[WL], (k_lift_level<WL, LIFT_TW, FRAME_BAND>), grid_r, LIFT_THREADS,
	            (lift_smem_bytes<WL, LIFT_TW, FRAME_BAND>()), p);
	p.frame = FRAME_COLS;
	const dim3 grid_c(2, (p.th + LIFT_TH - 1) / LIFT_TH, p.channels * n_images);
	AKOD_BYTES(c, (uint64_t)4 * 4 * FRAME_BAND * LIFT_TH * grid_c.x * grid_c.y * grid_c.z);
	AKOD_LAUNCH(c, names[WL], (k_lift_level<WL, FRAME_BAND, LIFT_TH>), grid_c, LIFT_THREADS,
	            (lift_smem_bytes<WL, FRAME_BAND, LIFT_TH>()), p);
	return AKOD_OK;
}

// strip kernel: pick the rows-per-CTA so that the grid fills the machine a few times over while the
// warm-up rows (2*LAT per CTA) stay a small fraction
// Rows per CTA for the strip kernels. A CTA marches 8k - 2*LAT output rows after 2*LAT warm-up rows, so larger k
// means less redundant work but fewer CTAs. Pick the k whose grid wastes least: (last-wave occupancy) x (useful
// fraction of the rows a CTA touches). 'slots' = CTAs resident on the whole GPU at once.
static uint32_t strip_split(uint32_t rows, uint64_t ctas_per_row_split, uint64_t slots, int lat)
{
	uint32_t best = 0;
	double best_eff = -1.0;
	for (uint32_t k = 32; k >= 4; k >>= 1)
	{
		const uint32_t split = 8 * k - 2 * (uint32_t)lat;
		const uint64_t ctas = ctas_per_row_split * ((rows + split - 1) / split);
		const uint64_t waves = (ctas + slots - 1) / slots;
		const double eff = ((double)ctas / (double)(waves * slots)) * ((double)split / (double)(split + 2 * lat));
		if (eff > best_eff + 1e-9)
		{
			best_eff = eff;
			best = split;
		}
	}
	return best;
}

template <int WL>
static int launch_lift_strip(akodContext* c, const LiftParams& p, uint32_t n_images)
{
	constexpr int LAT = StripGeom<WL>::LAT;
	const uint32_t strips = (p.tw + FS_TW - 1) / FS_TW;
	const uint32_t split = strip_split(p.th, (uint64_t)strips * p.channels * n_images, (uint64_t)c->sm_count * 6, LAT);
	StripParams sp;
	sp.p = p;
	sp.split = split;
	bool plain = true, gate = false;
	for (uint32_t ch = 0; ch < p.channels; ch++)
	{
		plain = plain && p.q[ch] <= 1 && p.g[ch] == 0;
		gate = gate || p.g[ch] >= p.q[ch];
	}
	const dim3 grid(strips, (p.th + split - 1) / split, p.channels * n_images);
	static const char* const names[3] = {"lift_strip_dd137", "lift_strip_cdf53", "lift_strip_haar"};
	AKOD_BYTES(c, (uint64_t)4 * p.cw * p.ch * p.channels * n_images); // every sample read once, every coefficient written once
	if (plain)
		AKOD_LAUNCH(c, names[WL], (k_lift_strip<WL, FS_PLAIN>), grid, FS_THREADS, 0, sp);
	else if (!gate)
		AKOD_LAUNCH(c, names[WL], (k_lift_strip<WL, FS_QUANT>), grid, FS_THREADS, 0, sp);
	else
		AKOD_LAUNCH(c, names[WL], (k_lift_strip<WL, FS_GATE>), grid, FS_THREADS, 0, sp);
	return AKOD_OK;
}

// level 0 straight from the interleaved RGBA8 image (lift_strip4.cuh)
template <int WL>
static int launch_lift_strip4(akodContext* c, const LiftParams& p, uint32_t n_images, const uint8_t* rgba, uint64_t rgba_is,
                              uint64_t rgba_rs, int color, int discard)
{
	constexpr int LAT = StripGeom<WL>::LAT;
	if (!(c->strip4_attr_done & (1u << WL)))
	{
		AKOD_TRY(cudaSetDevice(c->device));
		AKOD_TRY(cudaFuncSetAttribute(k_lift_strip4<WL, FS_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F4_SMEM));
		AKOD_TRY(cudaFuncSetAttribute(k_lift_strip4<WL, FS_QUANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F4_SMEM));
		AKOD_TRY(cudaFuncSetAttribute(k_lift_strip4<WL, FS_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F4_SMEM));
		c->strip4_attr_done |= 1u << WL;
	}
	const uint32_t strips = (p.tw + F4_TW - 1) / F4_TW;
	const uint32_t split = strip_split(p.th, (uint64_t)strips * n_images, (uint64_t)c->sm_count * F4_CTAS, LAT);
	Strip4Params sp;
	sp.p = p;
	sp.rgba = rgba;
	sp.rgba_is = rgba_is;
	sp.rgba_rs = (uint32_t)rgba_rs;
	sp.color = color;
	sp.discard = discard;
	sp.split = split;
	bool plain = true, gate = false;
	for (uint32_t ch = 0; ch < p.channels; ch++)
	{
		plain = plain && p.q[ch] <= 1 && p.g[ch] == 0;
		gate = gate || p.g[ch] >= p.q[ch];
	}
	const dim3 grid(strips, (p.th + split - 1) / split, n_images);
	static const char* const names[3] = {"lift_strip4_dd137", "lift_strip4_cdf53", "lift_strip4_haar"};
	AKOD_BYTES(c, (uint64_t)3 * p.cw * p.ch * p.channels * n_images); // every u8 sample read once, every int16 coefficient written once
	if (plain)
		AKOD_LAUNCH(c, names[WL], (k_lift_strip4<WL, FS_PLAIN>), grid, F4_THREADS, F4_SMEM, sp);
	else if (!gate)
		AKOD_LAUNCH(c, names[WL], (k_lift_strip4<WL, FS_QUANT>), grid, F4_THREADS, F4_SMEM, sp);
	else
		AKOD_LAUNCH(c, names[WL], (k_lift_strip4<WL, FS_GATE>), grid, F4_THREADS, F4_SMEM, sp);
	return AKOD_OK;
}

template <int WL>
static int launch_unlift_strip(akodContext* c, const UnliftParams& p, uint32_t n_images, bool v1)
{
	constexpr int LAT = StripGeom<WL>::LAT;
	const uint32_t width = v1 ? US_TW : UT_TW;
	const uint32_t strips = (p.hw + width - 1) / width;
	const uint32_t split = strip_split(p.hh, (uint64_t)strips * p.channels * n_images, (uint64_t)c->sm_count * 6, LAT);
	UnstripParams up;
	up.p = p;
	up.split = split;
	const dim3 grid(strips, (p.hh + split - 1) / split, p.channels * n_images);
	static const char* const names[3] = {"unlift_strip_dd137", "unlift_strip_cdf53", "unlift_strip_haar"};
	static const char* const names_v1[3] = {"unlift_strip_v1_dd137", "unlift_strip_v1_cdf53", "unlift_strip_v1_haar"};
	AKOD_BYTES(c, (uint64_t)4 * p.tw * p.th * p.channels * n_images);
	if (v1)
		AKOD_LAUNCH(c, names_v1[WL], k_unlift_strip_v1<WL>, grid, US_THREADS, 0, up);
	else
		AKOD_LAUNCH(c, names[WL], k_unlift_strip<WL>, grid, UT_THREADS, 0, up);
	return AKOD_OK;
}

template <int WL>
static int launch_unlift_level(akodContext* c, const UnliftParams& p, uint32_t n_images)
{
	const dim3 grid((p.hw + LIFT_TW - 1) / LIFT_TW, (p.hh + LIFT_TH - 1) / LIFT_TH, p.channels * n_images);
	static const char* const names[3] = {"unlift_dd137", "unlift_cdf53", "unlift_haar"};
	AKOD_BYTES(c, (uint64_t)4 * p.tw * p.th * p.channels * n_images);
	AKOD_LAUNCH(c, names[WL], k_unlift_level<WL>, grid, LIFT_THREADS, unlift_smem_bytes<WL>(), p);
	return AKOD_OK;
}

template <int WL>
static int launch_unlift_frame(akodContext* c, const UnliftParams& p_in, uint32_t n_images)
{
	UnliftParams p = p_in;
	static const char* const names[3] = {"unlift_frame_dd137", "unlift_frame_cdf53", "unlift_frame_haar"};
	p.frame = FRAME_ROWS;
	const dim3 grid_r((p.hw + LIFT_TW - 1) / LIFT_TW, 2, p.channels * n_images);
	AKOD_BYTES(c, (uint64_t)4 * 4 * LIFT_TW * FRAME_BAND * grid_r.x * grid_r.y * grid_r.z);
	AKOD_LAUNCH(c, names[WL], (k_unlift_level<WL, LIFT_TW, FRAME_BAND>), grid_r, LIFT_THREADS,
	            (unlift_smem_bytes<WL, LIFT_TW, FRAME_BAND>()), p);
	p.frame = FRAME_COLS;
	const dim3 grid_c(2, (p.hh + LIFT_TH - 1) / LIFT_TH, p.channels * n_images);
	AKOD_BYTES(c, (uint64_t)4 * 4 * FRAME_BAND * LIFT_TH * grid_c.x * grid_c.y * grid_c.z);
	AKOD_LAUNCH(c, names[WL], (k_unlift_level<WL, FRAME_BAND, LIFT_TH>), grid_c, LIFT_THREADS,
	            (unlift_smem_bytes<WL, FRAME_BAND, LIFT_TH>()), p);
	return AKOD_OK;
}

// the tail of the pyramid (levels l0 .. levels-1) in one launch; see lift_small.cuh
static int launch_small(akodContext* c, const akodPlan* plan, uint32_t l0, bool forward, int16_t* planes, uint32_t planes_rs,
                        uint64_t planes_ps, uint64_t planes_is, int16_t* stream, uint64_t stream_is, uint32_t n_images)
{
	if (!c->small_attr_done)
	{
		const size_t most = sizeof(int16_t) * 2 * (size_t)SM_CAP;
		AKOD_TRY(cudaSetDevice(c->device));
		AKOD_TRY(cudaFuncSetAttribute(k_lift_small<SM_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
		AKOD_TRY(cudaFuncSetAttribute(k_unlift_small<SM_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
		AKOD_TRY(cudaFuncSetAttribute(k_lift_small<SM_THREADS_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
		AKOD_TRY(cudaFuncSetAttribute(k_unlift_small<SM_THREADS_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most));
		c->small_attr_done = true;
	}
	// shared memory by need: the planes of a tile batch are small and many, several CTAs then share an SM
	const uint32_t cap = (uint32_t)small_capacity(plan->level[l0].cw, plan->level[l0].ch, plan->levels - l0);
	const size_t smem = sizeof(int16_t) * 2 * (size_t)cap;
	SmallParams sp;
	memset(&sp, 0, sizeof(sp));
	sp.planes = planes;
	sp.planes_rs = planes_rs;
	sp.planes_ps = planes_ps;
	sp.planes_is = planes_is;
	sp.stream = stream;
	sp.stream_is = stream_is;
	sp.cw0 = plan->level[l0].cw;
	sp.ch0 = plan->level[l0].ch;
	sp.levels = plan->levels - l0;
	sp.channels = plan->channels;
	sp.cap = cap;
	sp.wrap = plan->wrap;
	sp.wavelet = plan->wavelet;
	const uint32_t c1 = plan->channels > 1 ? 1 : 0;
	for (uint32_t s = 0; s < sp.levels; s++)
	{
		const akodLevel* L = &plan->level[l0 + s];
		sp.lq[s].qy = L->q[0];
		sp.lq[s].qc = L->q[c1];
		sp.lq[s].gy = L->g[0];
		sp.lq[s].gc = L->g[c1];
	}
	{
		uint64_t samples = 0;
		for (uint32_t s = 0; s < sp.levels; s++)
			samples += (uint64_t)plan->level[l0 + s].cw * plan->level[l0 + s].ch;
		AKOD_BYTES(c, 4 * samples * plan->channels * n_images);
	}
	const bool tile_sized = (uint64_t)sp.cw0 * sp.ch0 <= 8192 && n_images > 1;
	if (forward && tile_sized)
		AKOD_LAUNCH(c, "lift_small", k_lift_small<SM_THREADS_TILE>, plan->channels * n_images, SM_THREADS_TILE, smem, sp);
	else if (forward)
		AKOD_LAUNCH(c, "lift_small", k_lift_small<SM_THREADS>, plan->channels * n_images, SM_THREADS, smem, sp);
	else if (tile_sized)
		AKOD_LAUNCH(c, "unlift_small", k_unlift_small<SM_THREADS_TILE>, plan->channels * n_images, SM_THREADS_TILE, smem, sp);
	else
		AKOD_LAUNCH(c, "unlift_small", k_unlift_small<SM_THREADS>, plan->channels * n_images, SM_THREADS, smem, sp);
	return AKOD_OK;
}

// The wrap mode a level's kernels are given. The four modes only differ in what an out-of-range tap reads
// (oracle/ako_oracle.c s_map, pinned against wavelet-*.c): Haar has no such tap, and for CDF 5/3 MIRROR maps indices
// like CLAMP (its two tap substitutions exist only in DD 13/7, wavelet-dd137.c:123, :164). Those levels take the
// CLAMP kernels -- which include the strip kernels -- with identical results.
static inline int akod_level_wrap(int level_wavelet, int wrap)
{
	if (level_wavelet == AKOD_HAAR || (level_wavelet == AKOD_CDF53 && wrap == AKOD_WRAP_MIRROR))
		return AKOD_WRAP_CLAMP;
	return wrap;
}

template <class P>
static inline P as_clamp(P p)
{
	p.wrap = AKOD_WRAP_CLAMP;
	return p;
}

static inline uint32_t akod_pad8(uint32_t v)
{
	return (v + 7u) & ~7u;
}

// Does the single-launch tail kernel (one CTA per plane, lift_small.cuh) take the pyramid from this level down? It is
// there for the latency of a lone image's small levels; a batch of many planes (tiles of a large image, many images)
// has the parallelism for the strip kernels, which run several times faster per sample, down to their smallest level.
static inline bool akod_small_from(const akodPlan* plan, uint32_t l, uint32_t n_members)
{
	const akodLevel* L = &plan->level[l];
	if (!small_eligible(L->cw, L->ch, plan->levels - l))
		return false;
	const bool strip_size = plan->wrap == AKOD_WRAP_CLAMP && L->cw >= 64 && L->ch >= 16;
	return !(strip_size && (uint64_t)n_members * plan->channels >= 128);
}

// u8: when not NULL, level 0 may read the interleaved RGBA8 image itself (the colour/format pass fused into the
// lifting kernel); *fused tells whether it did -- if not, the caller's planes must hold the formatted image.
struct LiftRgba
{
	const uint8_t* rgba;
	uint64_t rgba_is, rgba_rs;
	int color, discard;
};

static int lift_pyramid(akodContext* c, const akodPlan* plan, int16_t* d_planes, int16_t* d_scratch, int16_t* d_stream,
                        const akodBatch* b, const LiftRgba* u8, bool probe_only, bool* fused)
{
	if (fused)
		*fused = false;
	const uint32_t n = b ? b->n : 1;
	const uint64_t planes_is = b ? b->planes_stride : 0, scratch_is = b ? b->scratch_stride : 0;
	const uint64_t stream_is = b ? b->stream_stride : 0;

	// ping-pong: level l reads 'src' (cw x ch planes, rows src_rs apart) and writes the next LL into 'dst'. Rows of
	// the intermediate planes are padded to a multiple of 8 elements, so that every row starts on a 16-byte boundary
	// whatever the width (the strip kernels fetch rows with the TMA engine); the caller says how the level-0 planes
	// are laid out (akodBatch.planes_pitch, dense when 0).
	int16_t* src = d_planes;
	int16_t* dst = d_scratch;
	uint64_t src_is = planes_is, dst_is = scratch_is;
	uint32_t src_rs = (b && b->planes_pitch) ? b->planes_pitch : plan->w;
	uint64_t src_ps = (uint64_t)src_rs * plan->h;

	static const bool no_small = getenv("AKO_B200_NO_SMALL") != nullptr;
	for (uint32_t l = 0; l < plan->levels; l++)
	{
		const akodLevel* L = &plan->level[l];
		if (!no_small && akod_small_from(plan, l, n))
			return launch_small(c, plan, l, true, src, src_rs, src_ps, src_is, d_stream, stream_is, n);
		LiftParams p;
		memset(&p, 0, sizeof(p));
		p.in = src;
		p.in_rs = src_rs;
		p.in_ps = src_ps;
		p.in_is = src_is;
		p.cw = L->cw;
		p.ch = L->ch;
		p.tw = L->tw;
		p.th = L->th;
		p.wrap = akod_level_wrap(L->wavelet, plan->wrap);
		p.channels = plan->channels;
		p.stream = d_stream;
		p.stream_is = stream_is;
		if (l + 1 == plan->levels)
		{
			// coarsest level: its lowpass IS the stream's LP section (lifting.c:280-291)
			p.ll = d_stream + plan->off_lp[0];
			p.ll_rs = L->tw;
			p.ll_ps = (uint64_t)plan->lp_w * plan->lp_h;
			p.ll_is = stream_is;
		}
		else
		{
			p.ll = dst;
			p.ll_rs = akod_pad8(L->tw);
			p.ll_ps = (uint64_t)p.ll_rs * L->th;
			p.ll_is = dst_is;
		}
		for (uint32_t ch = 0; ch < plan->channels; ch++)
		{
			p.off_c[ch] = L->off_c[ch];
			p.q[ch] = L->q[ch] < 1 ? 1 : L->q[ch];
			p.g[ch] = L->g[ch];
			p.qmagic[ch] = (p.q[ch] > 1) ? (uint32_t)((((uint64_t)1 << 32) + p.q[ch] - 1) / (uint64_t)p.q[ch]) : 0;
			int c2 = 0;
			while ((1 << c2) < p.q[ch])
				c2++;
			p.qshift[ch] = 15 + c2;
			p.qmul[ch] = (uint32_t)((((uint64_t)1 << p.qshift[ch]) / (uint64_t)p.q[ch]) + 1);
		}
		int rc;
		static const bool no_strip = getenv("AKO_B200_NO_STRIP") != nullptr;
		static const bool no_fuse = getenv("AKO_B200_NO_FUSE") != nullptr;
		static const bool no_frame = getenv("AKO_B200_NO_FRAME") != nullptr;
		// another wrap mode than CLAMP: the strip kernels as CLAMP, then the frame of edge tiles again (launch_lift_frame)
		const bool framed = !no_strip && !no_frame && p.wrap != AKOD_WRAP_CLAMP && frame_worth(p.tw, p.th);
		const LiftParams ps = framed ? as_clamp(p) : p; // what the strip kernels are asked
		if (l == 0 && fused)
		{
			*fused = u8 != nullptr && !no_strip && !no_fuse &&
			         lift_strip4_eligible(ps, u8->rgba, u8->rgba_is, u8->rgba_rs);
			if (probe_only)
				return AKOD_OK;
		}
		if (l == 0 && fused && *fused)
		{
			if (L->wavelet == AKOD_DD137)
				rc = launch_lift_strip4<AKOD_DD137>(c, ps, n, u8->rgba, u8->rgba_is, u8->rgba_rs, u8->color, u8->discard);
			else if (L->wavelet == AKOD_CDF53)
				rc = launch_lift_strip4<AKOD_CDF53>(c, ps, n, u8->rgba, u8->rgba_is, u8->rgba_rs, u8->color, u8->discard);
			else
				rc = launch_lift_strip4<AKOD_HAAR>(c, ps, n, u8->rgba, u8->rgba_is, u8->rgba_rs, u8->color, u8->discard);
			if (rc == AKOD_OK && framed)
			{
				// the planes were never written: the frame converts the pixels it needs itself
				LiftParams pf = p;
				pf.rgba = u8->rgba;
				pf.rgba_is = u8->rgba_is;
				pf.rgba_rs = u8->rgba_rs;
				pf.rgba_color = u8->color;
				pf.rgba_discard = u8->discard;
				rc = (L->wavelet == AKOD_DD137) ? launch_lift_frame<AKOD_DD137>(c, pf, n) : launch_lift_frame<AKOD_CDF53>(c, pf, n);
			}
		}
		else if (!no_strip && lift_strip_eligible(p))
		{
			if (L->wavelet == AKOD_DD137)
				rc = launch_lift_strip<AKOD_DD137>(c, p, n);
			else if (L->wavelet == AKOD_CDF53)
				rc = launch_lift_strip<AKOD_CDF53>(c, p, n);
			else
				rc = launch_lift_strip<AKOD_HAAR>(c, p, n);
		}
		else if (framed && lift_strip_eligible(ps))
		{
			// (Haar never gets here: akod_level_wrap)
			if (L->wavelet == AKOD_DD137)
				rc = launch_lift_strip<AKOD_DD137>(c, ps, n);
			else
				rc = launch_lift_strip<AKOD_CDF53>(c, ps, n);
			if (rc == AKOD_OK)
				rc = (L->wavelet == AKOD_DD137) ? launch_lift_frame<AKOD_DD137>(c, p, n) : launch_lift_frame<AKOD_CDF53>(c, p, n);
		}
		else if (L->wavelet == AKOD_DD137)
			rc = launch_lift_level<AKOD_DD137>(c, p, n);
		else if (L->wavelet == AKOD_CDF53)
			rc = launch_lift_level<AKOD_CDF53>(c, p, n);
		else
			rc = launch_lift_level<AKOD_HAAR>(c, p, n);
		if (rc != AKOD_OK)
			return rc;

		// swap
		int16_t* t = src;
		src = dst;
		dst = t;
		const uint64_t ti = src_is;
		src_is = dst_is;
		dst_is = ti;
		src_rs = p.ll_rs;
		src_ps = p.ll_ps;
	}

	if (plan->levels == 0)
	{
		// w <= 2 or h <= 2: no lift at all; the planes are the LP section. (Outside the reference's own
		// well-defined domain; we define it as the dense copy.)
		for (uint32_t i = 0; i < n; i++)
			AKOD_TRY(cudaMemcpyAsync(d_stream + stream_is * i, d_planes + planes_is * i,
			                         sizeof(int16_t) * plan->stream_len, cudaMemcpyDeviceToDevice, c->stream));
	}
	return AKOD_OK;
}

extern "C" int akod_lift(akodContext* c, const akodPlan* plan, int16_t* d_planes, int16_t* d_scratch, int16_t* d_stream,
                         const akodBatch* b)
{
	akod_use(c);
	return lift_pyramid(c, plan, d_planes, d_scratch, d_stream, b, nullptr, false, nullptr);
}

// Will akod_format_lift run the colour/format pass inside the level-0 lifting kernel? (4-channel CLAMP images whose
// level 0 takes the strip kernel: lift_strip4.cuh.) The host asks so that its FORMAT / WAVELET events bracket what
// they name.
extern "C" int akod_format_lift_fuses(akodContext* c, uint32_t channels, uint32_t w, uint32_t h, uint64_t in_stride_px,
                                      const uint8_t* d_in, const akodPlan* plan, int16_t* d_planes, int16_t* d_scratch,
                                      int16_t* d_stream, const akodBatch* b)
{
	if (channels != 4 || (b && b->n_real) || plan->w != w || plan->h != h || plan->levels == 0)
		return 0;
	LiftRgba u8;
	u8.rgba = d_in;
	u8.rgba_is = b ? b->in_stride : 0;
	u8.rgba_rs = in_stride_px * 4;
	u8.color = u8.discard = 0;
	bool fused = false;
	if (lift_pyramid(c, plan, d_planes, d_scratch, d_stream, b, &u8, true, &fused) != AKOD_OK)
		return 0;
	return fused ? 1 : 0;
}

// akod_format_forward + akod_lift; when akod_format_lift_fuses() says so, level 0 reads the RGBA8 image itself and
// d_planes is only a ping-pong buffer of the coarser levels
extern "C" int akod_format_lift(akodContext* c, int discard, int color, uint32_t channels, uint32_t w, uint32_t h,
                                uint64_t in_stride_px, const uint8_t* d_in, const akodPlan* plan, int16_t* d_planes,
                                int16_t* d_scratch, int16_t* d_stream, const akodBatch* b)
{
	akod_use(c);
	if (akod_format_lift_fuses(c, channels, w, h, in_stride_px, d_in, plan, d_planes, d_scratch, d_stream, b))
	{
		LiftRgba u8;
		u8.rgba = d_in;
		u8.rgba_is = b ? b->in_stride : 0;
		u8.rgba_rs = in_stride_px * 4;
		u8.color = color;
		u8.discard = discard;
		bool fused = false;
		return lift_pyramid(c, plan, d_planes, d_scratch, d_stream, b, &u8, false, &fused);
	}
	const int rc = akod_format_forward(c, discard, color, channels, w, h, in_stride_px, d_in, d_planes, b);
	if (rc != AKOD_OK)
		return rc;
	return lift_pyramid(c, plan, d_planes, d_scratch, d_stream, b, nullptr, false, nullptr);
}

extern "C" int akod_unlift(akodContext* c, const akodPlan* plan, const int16_t* d_stream, int16_t* d_planes,
                           int16_t* d_scratch, const akodBatch* b)
{
	akod_use(c);
	const uint32_t n = b ? b->n : 1;
	const uint64_t planes_is = b ? b->planes_stride : 0, scratch_is = b ? b->scratch_stride : 0;
	const uint64_t stream_is = b ? b->stream_stride : 0;

	if (plan->levels == 0)
	{
		for (uint32_t i = 0; i < n; i++)
			AKOD_TRY(cudaMemcpyAsync(d_planes + planes_is * i, d_stream + stream_is * i,
			                         sizeof(int16_t) * plan->stream_len, cudaMemcpyDeviceToDevice, c->stream));
		return AKOD_OK;
	}

	// The finest level must land in d_planes; alternate buffers backwards from there.
	// level index l (0 = finest) writes to planes if l is even, scratch if odd. Rows of the intermediate planes are
	// padded to a multiple of 8 elements (see lift_pyramid); the level-0 planes as the caller says.
	const uint32_t pitch0 = (b && b->planes_pitch) ? b->planes_pitch : plan->w;
	static const bool no_small = getenv("AKO_B200_NO_SMALL") != nullptr;
	uint32_t l_top = plan->levels; // levels l_top .. levels-1 are done by the small kernel
	if (!no_small)
		for (uint32_t l = 0; l < plan->levels; l++)
			if (akod_small_from(plan, l, n))
			{
				l_top = l;
				break;
			}
	if (l_top < plan->levels)
	{
		const akodLevel* L = &plan->level[l_top];
		const bool to_planes = (l_top % 2) == 0;
		const uint32_t rs = (l_top == 0) ? pitch0 : akod_pad8(L->cw);
		int rc = launch_small(c, plan, l_top, false, to_planes ? d_planes : d_scratch, rs, (uint64_t)rs * L->ch,
		                      to_planes ? planes_is : scratch_is, const_cast<int16_t*>(d_stream), stream_is, n);
		if (rc != AKOD_OK)
			return rc;
	}
	for (uint32_t l = l_top; l-- > 0;)
	{
		const akodLevel* L = &plan->level[l];
		UnliftParams p;
		memset(&p, 0, sizeof(p));
		p.hw = L->tw;
		p.hh = L->th;
		p.tw = L->cw;
		p.th = L->ch;
		p.wrap = akod_level_wrap(L->wavelet, plan->wrap);
		p.channels = plan->channels;
		p.stream = d_stream;
		p.stream_is = stream_is;
		if (l + 1 == plan->levels)
		{
			p.ll = d_stream + plan->off_lp[0];
			p.ll_rs = plan->lp_w;
			p.ll_ps = (uint64_t)plan->lp_w * plan->lp_h;
			p.ll_is = stream_is;
		}
		else
		{
			const bool from_planes = ((l + 1) % 2) == 0;
			p.ll = from_planes ? d_planes : d_scratch;
			p.ll_rs = akod_pad8(L->tw);
			p.ll_ps = (uint64_t)p.ll_rs * L->th;
			p.ll_is = from_planes ? planes_is : scratch_is;
		}
		const bool to_planes = (l % 2) == 0;
		p.out = to_planes ? d_planes : d_scratch;
		p.out_rs = (l == 0) ? pitch0 : akod_pad8(L->cw);
		p.out_ps = (uint64_t)p.out_rs * L->ch;
		p.out_is = to_planes ? planes_is : scratch_is;
		for (uint32_t ch = 0; ch < plan->channels; ch++)
			p.off_c[ch] = L->off_c[ch];
		int rc;
		static const bool no_strip = getenv("AKO_B200_NO_STRIP") != nullptr;
		static const bool no_frame = getenv("AKO_B200_NO_FRAME") != nullptr;
		const bool framed = !no_strip && !no_frame && p.wrap != AKOD_WRAP_CLAMP &&
		                    frame_worth(p.hw, p.hh);
		const UnliftParams ps = framed ? as_clamp(p) : p; // what the strip kernels are asked
		const bool v2 = !no_strip && unlift_strip_eligible(ps), v1 = !no_strip && !v2 && unlift_strip_v1_eligible(ps);
		if (v2 || v1)
		{
			if (L->wavelet == AKOD_DD137)
				rc = launch_unlift_strip<AKOD_DD137>(c, ps, n, v1);
			else if (L->wavelet == AKOD_CDF53)
				rc = launch_unlift_strip<AKOD_CDF53>(c, ps, n, v1);
			else
				rc = launch_unlift_strip<AKOD_HAAR>(c, ps, n, v1);
			if (rc == AKOD_OK && framed)
				rc = (L->wavelet == AKOD_DD137) ? launch_unlift_frame<AKOD_DD137>(c, p, n) : launch_unlift_frame<AKOD_CDF53>(c, p, n);
		}
		else if (L->wavelet == AKOD_DD137)
			rc = launch_unlift_level<AKOD_DD137>(c, p, n);
		else if (L->wavelet == AKOD_CDF53)
			rc = launch_unlift_level<AKOD_CDF53>(c, p, n);
		else
			rc = launch_unlift_level<AKOD_HAAR>(c, p, n);
		if (rc != AKOD_OK)
			return rc;
	}
	return AKOD_OK;
}

// ------------------------------------------------------------------------------------------------
// Kagari

extern "C" int akod_kagari_encode(akodContext* c, uint64_t n_values, const int16_t* d_in, uint64_t in_stride,
                                  uint8_t* d_out, uint64_t out_stride, uint64_t out_cap, uint64_t* d_bits,
                                  uint32_t n_images)
{
	akod_use(c);
	if (n_values == 0 || n_values >= ((uint64_t)1 << 32))
		return AKOD_ERROR;
	const uint32_t nblocks = (uint32_t)((n_values + KG_BLOCK - 1) / KG_BLOCK);
	const uint64_t per_img = nblocks;

	void* ws;
	// The single-pass chained-scan encoder (k_kg_fused) is bit-exact but measured slower than the three passes on
	// B200 (0.97 ms vs 0.54 ms for 8 x 16 M values): its two dependent look-backs put ~4 global round trips on every
	// block's critical path. Kept selectable for experiments.
	static const bool fused = getenv("AKO_B200_KG_FUSED") != nullptr;
	if (fused)
	{
		// single pass: [bit state 16 B | run state 8 B] per block, one ticket per image; all zero at launch
		const size_t blocks_all = (size_t)per_img * n_images;
		const size_t need1 = blocks_all * (sizeof(KgBitState) + sizeof(unsigned long long)) + sizeof(uint32_t) * n_images + 64;
		int rc1 = akod_workspace(c, AKOD_WS_KAGARI, need1, &ws);
		if (rc1 != AKOD_OK)
			return rc1;
		AKOD_TRY(cudaMemsetAsync(ws, 0, need1, c->stream));
		KgBitState* bit_state = (KgBitState*)ws;
		unsigned long long* run_state = (unsigned long long*)(bit_state + blocks_all);
		uint32_t* ticket = (uint32_t*)(run_state + blocks_all);
		const dim3 grid1(nblocks, n_images);
		AKOD_LAUNCH(c, "kagari_encode", k_kg_fused, grid1, KG_THREADS, 0, d_in, in_stride, n_values, run_state, bit_state, ticket,
		            nblocks, d_out, out_stride, out_cap * 8, d_bits);
		return AKOD_OK;
	}
	const size_t need = (size_t)per_img * n_images * (sizeof(uint64_t) + 3 * sizeof(uint32_t) + sizeof(uint32_t) * KG_SLOT_WORDS + 1) + 64;
	int rc = akod_workspace(c, AKOD_WS_KAGARI, need, &ws);
	if (rc != AKOD_OK)
		return rc;
	uint64_t* blk_off = (uint64_t*)ws;
	uint32_t* blk_start = (uint32_t*)(blk_off + per_img * n_images);
	uint32_t* blk_bits = blk_start + per_img * n_images;
	uint32_t* slots = blk_bits + per_img * n_images; // KG_SLOT_WORDS per block
	uint32_t* blk_own = slots + per_img * n_images * KG_SLOT_WORDS;
	uint8_t* blk_first = (uint8_t*)(blk_own + per_img * n_images);

	// blocks per warp: four when that still gives every SM its five CTAs, fewer for a small batch (one image alone was
	// 246 CTAs of 32 blocks each: a latency chain on 1.7 CTAs per SM)
	uint32_t bpw = KGL_BLOCKS_PER_WARP;
	while (bpw > 1 && (uint64_t)((nblocks + KGL_WARPS * bpw - 1) / (KGL_WARPS * bpw)) * n_images < (uint64_t)c->sm_count * 5)
		bpw >>= 1;
	const dim3 lgrid((nblocks + KGL_WARPS * bpw - 1) / (KGL_WARPS * bpw), n_images);
	AKOD_BYTES(c, 2 * n_values * n_images);
	AKOD_LAUNCH(c, "kagari_starts", k_kg_starts, dim3((nblocks + KG_STARTS_PER_CTA - 1) / KG_STARTS_PER_CTA, n_images), KG_THREADS,
	            0, d_in, in_stride, n_values, blk_own, blk_first, nblocks);
	AKOD_LAUNCH(c, "kagari_scan_max", k_kg_scan_max, n_images, 1024, 0, blk_own, blk_start, nblocks);
	AKOD_BYTES(c, 2 * n_values * n_images);
	AKOD_LAUNCH(c, "kagari_lengths", k_kg_lengths, lgrid, KG_THREADS, 0, d_in, in_stride, n_values, blk_start, blk_bits,
	            nblocks, slots, blk_own, blk_first, bpw);
	AKOD_LAUNCH(c, "kagari_scan_sum", k_kg_scan_sum, n_images, 1024, 0, blk_bits, blk_off, nblocks, d_bits);
	const dim3 zgrid((nblocks + 255) / 256, n_images);
	AKOD_LAUNCH(c, "kagari_zero_edges", k_kg_zero_edges, zgrid, 256, 0, blk_off, blk_bits, nblocks, d_out, out_stride,
	            out_cap * 8);
	// algorithmic bytes of pass 3 = the blob bytes, known only to the caller once the sizes are read back
	AKOD_LAUNCH(c, "kagari_pack", k_kg_pack, dim3((nblocks + KG_PACK_PER_CTA - 1) / KG_PACK_PER_CTA, n_images), KG_THREADS, 0, d_in, in_stride, n_values, blk_start, blk_off, blk_bits,
	            nblocks, d_out, out_stride, out_cap * 8, slots);
	return AKOD_OK;
}

// passes 1 and 2 of the encoder only: d_bits[i] = exact bit length of image i's Kagari stream (ratio search probes)
extern "C" int akod_kagari_bits(akodContext* c, uint64_t n_values, const int16_t* d_in, uint64_t in_stride, uint64_t* d_bits,
                                uint32_t n_images)
{
	akod_use(c);
	if (n_values == 0 || n_values >= ((uint64_t)1 << 32))
		return AKOD_ERROR;
	const uint32_t nblocks = (uint32_t)((n_values + KG_BLOCK - 1) / KG_BLOCK);
	const uint64_t per_img = nblocks;
	void* ws;
	const size_t need = (size_t)per_img * n_images * (sizeof(uint64_t) + 3 * sizeof(uint32_t) + sizeof(uint32_t) * KG_SLOT_WORDS + 1) + 64;
	int rc = akod_workspace(c, AKOD_WS_KAGARI, need, &ws);
	if (rc != AKOD_OK)
		return rc;
	uint64_t* blk_off = (uint64_t*)ws;
	uint32_t* blk_start = (uint32_t*)(blk_off + per_img * n_images);
	uint32_t* blk_bits = blk_start + per_img * n_images;
	uint32_t* slots = blk_bits + per_img * n_images;
	uint32_t* blk_own = slots + per_img * n_images * KG_SLOT_WORDS;
	uint8_t* blk_first = (uint8_t*)(blk_own + per_img * n_images);
	// blocks per warp: four when that still gives every SM its five CTAs, fewer for a small batch (one image alone was
	// 246 CTAs of 32 blocks each: a latency chain on 1.7 CTAs per SM)
	uint32_t bpw = KGL_BLOCKS_PER_WARP;
	while (bpw > 1 && (uint64_t)((nblocks + KGL_WARPS * bpw - 1) / (KGL_WARPS * bpw)) * n_images < (uint64_t)c->sm_count * 5)
		bpw >>= 1;
	const dim3 lgrid((nblocks + KGL_WARPS * bpw - 1) / (KGL_WARPS * bpw), n_images);
	AKOD_BYTES(c, 2 * n_values * n_images);
	AKOD_LAUNCH(c, "kagari_starts", k_kg_starts, dim3((nblocks + KG_STARTS_PER_CTA - 1) / KG_STARTS_PER_CTA, n_images), KG_THREADS,
	            0, d_in, in_stride, n_values, blk_own, blk_first, nblocks);
	AKOD_LAUNCH(c, "kagari_scan_max", k_kg_scan_max, n_images, 1024, 0, blk_own, blk_start, nblocks);
	AKOD_BYTES(c, 2 * n_values * n_images);
	AKOD_LAUNCH(c, "kagari_lengths", k_kg_lengths, lgrid, KG_THREADS, 0, d_in, in_stride, n_values, blk_start, blk_bits,
	            nblocks, slots, blk_own, blk_first, bpw);
	AKOD_LAUNCH(c, "kagari_scan_sum", k_kg_scan_sum, n_images, 1024, 0, blk_bits, blk_off, nblocks, d_bits);
	return AKOD_OK;
}

// ------------------------------------------------------------------------------------------------
// ratio search support: quantise + gate an UNQUANTISED coefficient stream again with another schedule

struct RequantParams
{
	const int16_t* in;
	int16_t* out;
	uint64_t off_c; // channel 0's C subband; channel ch sits ch * (1 + 3*band) further
	uint32_t band;  // tw * th
	uint32_t channels;
	int16_t q[AKOD_MAX_CHANNELS], g[AKOD_MAX_CHANNELS];
	uint32_t qmagic[AKOD_MAX_CHANNELS];
};

__global__ void __launch_bounds__(256) k_requant(const RequantParams p)
{
	const uint32_t ch = blockIdx.y;
	const uint64_t base = p.off_c + (uint64_t)ch * (1 + 3ull * p.band);
	const int q = p.q[ch], g = p.g[ch];
	const uint32_t magic = p.qmagic[ch];
	if (blockIdx.x == 0 && threadIdx.x == 0)
		p.out[base - 1] = (int16_t)q; // akoLiftHead
	const uint64_t n = 3ull * p.band;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		p.out[base + i] = gate_quantize(p.in[base + i], q, g, magic);
}

// d_in: stream produced with q = 1, g = 0 on every level; d_out: what akod_lift would have produced with 'plan'
extern "C" int akod_requantize(akodContext* c, const akodPlan* plan, const int16_t* d_in, int16_t* d_out)
{
	akod_use(c);
	const uint64_t lp = (uint64_t)plan->lp_w * plan->lp_h * plan->channels;
	AKOD_TRY(cudaMemcpyAsync(d_out, d_in, sizeof(int16_t) * lp, cudaMemcpyDeviceToDevice, c->stream));
	for (uint32_t l = 0; l < plan->levels; l++)
	{
		const akodLevel* L = &plan->level[l];
		RequantParams p;
		memset(&p, 0, sizeof(p));
		p.in = d_in;
		p.out = d_out;
		p.off_c = L->off_c[0];
		p.band = L->tw * L->th;
		p.channels = plan->channels;
		for (uint32_t ch = 0; ch < plan->channels; ch++)
		{
			p.q[ch] = L->q[ch] < 1 ? 1 : L->q[ch];
			p.g[ch] = L->g[ch];
			p.qmagic[ch] = (p.q[ch] > 1) ? (uint32_t)((((uint64_t)1 << 32) + p.q[ch] - 1) / (uint64_t)p.q[ch]) : 0;
		}
		const uint64_t n = 3ull * p.band;
		const dim3 grid(akod_stream_grid(c, n, 256, 8), plan->channels);
		AKOD_BYTES(c, 4 * n * plan->channels);
		AKOD_LAUNCH(c, "requantize", k_requant, grid, 256, 0, p);
	}
	return AKOD_OK;
}

extern "C" int akod_kagari_decode(akodContext* c, uint64_t n_values, const uint8_t* d_in, const uint64_t* d_off,
                                  const uint64_t* d_size, uint64_t max_in_size, int16_t* d_out, uint64_t out_stride,
                                  uint64_t* d_result, uint32_t n_images)
{
	akod_use(c);
	if (n_values == 0 || n_images == 0 || n_values >= ((uint64_t)1 << 32))
		return AKOD_ERROR;
	static const bool force_sequential = getenv("AKO_B200_SEQ_DECODE") != nullptr;
	if (force_sequential)
	{
		AKOD_LAUNCH(c, "kagari_decode_seq", k_kd_sequential, n_images, 32, 0, d_in, d_off, d_size, n_values, d_out,
		            out_stride, d_result, (const KdImage*)nullptr, 0);
		return AKOD_OK;
	}

	// geometry shared by all images of the batch (sized for the largest block)
	const uint64_t max_bits = max_in_size * 8;
	const uint32_t nblk1 = (uint32_t)(max_bits / KD_CTA_BITS) + 1;

	// workspace carve-up
	size_t bytes = 0;
	auto carve = [&bytes](size_t n) {
		const size_t at = bytes;
		bytes += (n + 255) & ~(size_t)255;
		return at;
	};
	const size_t o_info = carve(sizeof(KdImage) * n_images);
	const size_t o_ends_a = carve(sizeof(uint64_t) * nblk1 * n_images);
	const size_t o_ends_b = carve(sizeof(uint64_t) * nblk1 * n_images);
	const size_t o_sub = carve(sizeof(KdSubState) * (size_t)nblk1 * KD_THREADS * n_images);
	// zeroed before every decode: look-back records, inclusive prefixes, CTA tickets, big-run counters
	const size_t o_look = carve(sizeof(KfLook) * (size_t)nblk1 * n_images);
	const size_t o_prefix = carve(sizeof(uint64_t) * (size_t)nblk1 * n_images);
	const size_t o_ticket = carve(sizeof(uint32_t) * n_images);
	const size_t o_bigc = carve(sizeof(uint32_t) * n_images);
	const size_t o_zero_end = bytes;
	const uint32_t big_cap = (uint32_t)(n_values / KT_BIG) + 16;
	const size_t o_big = carve(sizeof(KtRun) * (size_t)big_cap * n_images);
	void* ws;
	int rc = akod_workspace(c, AKOD_WS_KAGARI, bytes, &ws);
	if (rc != AKOD_OK)
		return rc;
	uint8_t* w8 = (uint8_t*)ws;
	KdImage* info = (KdImage*)(w8 + o_info);
	uint64_t* ends_a = (uint64_t*)(w8 + o_ends_a);
	uint64_t* ends_b = (uint64_t*)(w8 + o_ends_b);
	KdSubState* sub = (KdSubState*)(w8 + o_sub);
	KfLook* look = (KfLook*)(w8 + o_look);
	uint64_t* prefix = (uint64_t*)(w8 + o_prefix);
	uint32_t* ticket = (uint32_t*)(w8 + o_ticket);
	KtRun* big_list = (KtRun*)(w8 + o_big);
	uint32_t* big_count = (uint32_t*)(w8 + o_bigc);

	AKOD_TRY(cudaMemsetAsync(w8 + o_look, 0, o_zero_end - o_look, c->stream));
	AKOD_LAUNCH(c, "kagari_dec_init", k_kd_init, (n_images + 63) / 64, 64, 0, info, n_images);
	const dim3 grid1(nblk1, n_images);
	for (int run = 0; run < KD_MAX_RUNS; run++)
	{
		uint64_t* prev = (run & 1) ? ends_a : ends_b;
		uint64_t* next = (run & 1) ? ends_b : ends_a;
		AKOD_LAUNCH(c, "kagari_dec_sync", k_kd_sync, grid1, KD_THREADS, 0, d_in, d_off, d_size, nblk1, run, prev, next, sub, info);
	}
	static_assert((KD_MAX_RUNS & 1) == 0, "the last run writes ends_b");
	AKOD_BYTES(c, 2 * n_values * n_images); // the decoded values, written by this kernel and k_kt_fill together
	AKOD_LAUNCH(c, "kagari_dec_expand", k_kd_decode, grid1, KD_THREADS, 0, d_in, d_off, d_size, nblk1, sub, ends_b, look, prefix,
	            ticket, info, d_out, out_stride, n_values, big_list, big_count, big_cap);
	{
		// enough warps to saturate HBM with stores, whatever the number of big runs turns out to be
		const uint64_t pieces_max = (uint64_t)big_cap;
		const uint32_t want = (uint32_t)c->sm_count * 8;
		const uint32_t gx = (uint32_t)((pieces_max + 7) / 8 < want ? (pieces_max + 7) / 8 : want);
		const dim3 gridf(gx ? gx : 1, n_images);
		AKOD_LAUNCH(c, "kagari_dec_fill", k_kt_fill, gridf, 256, 0, big_list, big_count, big_cap, d_out, out_stride, info,
		            n_values, d_size, d_result);
	}
	// Blocks the parallel decoder does not accept (broken input, or no fixed point within KD_MAX_RUNS) are decoded by
	// one thread on the device exactly as the reference would; a no-op for every well-formed block.
	AKOD_LAUNCH(c, "kagari_dec_rescue", k_kd_sequential, n_images, 32, 0, d_in, d_off, d_size, n_values, d_out, out_stride,
	            d_result, (const KdImage*)info, 1);
	return AKOD_OK;
}

// ------------------------------------------------------------------------------------------------
// container

struct TileGrid
{
	akodTiles t;
};

__device__ __forceinline__ void tile_locate(const akodTiles& T, uint32_t t, uint32_t& g, uint32_t& k)
{
	const uint32_t ty = t / T.tiles_x, tx = t - ty * T.tiles_x;
	const uint32_t ex = tx >= T.full_x, ey = ty >= T.full_y;
	g = ex + 2 * ey;
	k = (ex && ey) ? 0 : ex ? ty : ey ? tx : ty * T.full_x + tx;
}

// One CTA per image: byte offset of every tile block inside the blob (an exclusive scan of 4 + block size over
// the tiles in raster order, compression.c:30-55 / encode.c:170-182) and the blob size. A tile that did not fit
// its capacity voids the image (total = 0). off is [n_images][n_tiles].
__global__ void __launch_bounds__(256)
    k_tile_offsets(const uint64_t* __restrict__ bits, const TileGrid G, int with_heads, uint64_t* __restrict__ off,
                   uint64_t* __restrict__ total)
{
	__shared__ uint64_t warp_sum[8];
	__shared__ uint64_t carry_sm;
	__shared__ int bad_sm;
	const akodTiles& T = G.t;
	const uint32_t img = blockIdx.x, n_tiles = T.tiles_x * T.tiles_y;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (threadIdx.x == 0)
	{
		carry_sm = 16;
		bad_sm = 0;
	}
	__syncthreads();
	for (uint32_t t0 = 0; t0 < n_tiles; t0 += 256)
	{
		const uint32_t t = t0 + threadIdx.x;
		uint64_t len = 0;
		if (t < n_tiles)
		{
			uint32_t g, k;
			tile_locate(T, t, g, k);
			const uint64_t bytes = (bits[T.bits_base[g] + (uint64_t)k * T.n_images + img] + 7) >> 3;
			if (bytes > T.group_cap[g])
				bad_sm = 1;
			len = bytes + (with_heads ? 4 : 0);
		}
		uint64_t incl = len;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const uint64_t o = __shfl_up_sync(AKOD_FULL_MASK, incl, d);
			if (lane >= d)
				incl += o;
		}
		if (lane == 31)
			warp_sum[wid] = incl;
		__syncthreads();
		uint64_t before = carry_sm;
		for (int w = 0; w < wid; w++)
			before += warp_sum[w];
		if (t < n_tiles)
			off[(uint64_t)img * n_tiles + t] = before + incl - len;
		__syncthreads();
		if (threadIdx.x == 255)
			carry_sm = before + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0)
		total[img] = bad_sm ? 0 : carry_sm;
}

struct HeadWords
{
	uint32_t w[4];
};

// blockIdx.x = (image * tiles + tile) * chunks + chunk
__global__ void __launch_bounds__(256)
    k_assemble(const HeadWords head, const uint8_t* __restrict__ blocks, const TileGrid G, const uint64_t* __restrict__ bits,
               const uint64_t* __restrict__ tile_off, const uint64_t* __restrict__ total, int with_heads,
               uint8_t* __restrict__ out, uint64_t out_stride, uint32_t chunks)
{
	const akodTiles& T = G.t;
	const uint32_t n_tiles = T.tiles_x * T.tiles_y;
	const uint32_t chunk = blockIdx.x % chunks;
	const uint32_t it = blockIdx.x / chunks;
	const uint32_t img = it / n_tiles, t = it - img * n_tiles;
	if (total[img] == 0)
		return;
	uint32_t g, k;
	tile_locate(T, t, g, k);
	const uint64_t member = (uint64_t)k * T.n_images + img;
	const uint64_t size = (bits[T.bits_base[g] + member] + 7) >> 3;
	out += out_stride * img;
	uint8_t* dst = out + tile_off[(uint64_t)img * n_tiles + t];
	if (t == 0 && chunk == 0 && threadIdx.x < 16)
		out[threadIdx.x] = (uint8_t)(head.w[threadIdx.x >> 2] >> (8 * (threadIdx.x & 3)));
	if (with_heads)
	{
		if (chunk == 0 && threadIdx.x < 4)
			dst[threadIdx.x] = (uint8_t)((uint32_t)size >> (8 * threadIdx.x)); // akoBlockHead, compression.c:30-33
		dst += 4;
	}
	const uint8_t* src = blocks + T.group_base[g] + member * T.group_stride[g];
	// The block region is 16-byte aligned, its place in the blob is not: every thread produces one 16-byte aligned
	// chunk of the destination from two aligned 16-byte loads, shifted by the byte misalignment.
	const uint32_t mis = (uint32_t)((uintptr_t)dst & 15);             // dst = dst_al + mis
	const uint64_t lead = mis ? min((uint64_t)(16 - mis), size) : 0;  // bytes before the first aligned chunk
	const uint64_t nchunks = (size - lead) / 16;
	const uint64_t stride = (uint64_t)chunks * blockDim.x, tid0 = (uint64_t)chunk * blockDim.x + threadIdx.x;
	if (tid0 < lead)
		dst[tid0] = src[tid0];
	const uint32_t sh = (uint32_t)(lead & 15);                        // chunk k reads src bytes [lead + 16k, +16): offset sh in aligned words
	const uint4* src4 = reinterpret_cast<const uint4*>(src);
	uint4* dst4 = reinterpret_cast<uint4*>(dst + lead);
	for (uint64_t kk = tid0; kk < nchunks; kk += stride)
	{
		const uint64_t s0 = (lead + 16 * kk) >> 4;
		const uint4 a = __ldg(src4 + s0);
		uint4 o = a;
		if (sh)
		{
			const uint4 b = __ldg(src4 + s0 + 1); // stays inside the 16-byte padded block region
			const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
			const uint32_t wo = sh >> 2, bs = 8 * (sh & 3);
			uint32_t r[4];
#pragma unroll
			for (int i = 0; i < 4; i++)
			{
				// bytes [sh + 4i, sh + 4i + 4) of the 32-byte window (little endian)
				uint32_t lo = 0, hi = 0;
#pragma unroll
				for (int j = 0; j < 8; j++)
				{
					lo = (j == (int)wo + i) ? w[j] : lo;
					hi = (j == (int)wo + i + 1) ? w[j] : hi;
				}
				r[i] = __funnelshift_r(lo, hi, bs);
			}
			o = make_uint4(r[0], r[1], r[2], r[3]);
		}
		dst4[kk] = o;
	}
	for (uint64_t i = lead + 16 * nchunks + tid0; i < size; i += stride)
		dst[i] = src[i];
}

extern "C" int akod_assemble(akodContext* c, const uint8_t head16[16], const akodTiles* tiles, const uint8_t* d_blocks,
                             const uint64_t* d_bits, int with_heads, uint8_t* d_out, uint64_t out_stride, uint64_t* d_total)
{
	akod_use(c);
	const uint64_t n_tiles = (uint64_t)tiles->tiles_x * tiles->tiles_y, n_images = tiles->n_images;
	void* ws;
	int rc = akod_workspace(c, AKOD_WS_KAGARI2, sizeof(uint64_t) * (size_t)n_tiles * n_images, &ws);
	if (rc != AKOD_OK)
		return rc;
	uint64_t* d_tile_off = (uint64_t*)ws;
	HeadWords hw;
	memcpy(hw.w, head16, 16);
	TileGrid G;
	G.t = *tiles;
	AKOD_LAUNCH(c, "tile_offsets", k_tile_offsets, (unsigned)n_images, 256, 0, d_bits, G, with_heads, d_tile_off, d_total);
	// CTAs per block: a lone large block is spread over the GPU, a crowd of blocks gets few CTAs of few threads each
	uint64_t largest = 0;
	for (int g = 0; g < 4; g++)
		largest = tiles->group_cap[g] > largest ? tiles->group_cap[g] : largest;
	unsigned chunks = (unsigned)c->sm_count, threads = 256;
	if (n_tiles * n_images >= 64)
		chunks = 4;
	if (largest <= (16u << 10))
	{
		chunks = 1;
		threads = 64;
	}
	if (n_tiles * n_images * chunks >= ((uint64_t)1 << 31))
		return AKOD_ERROR;
	AKOD_LAUNCH(c, "assemble", k_assemble, (unsigned)(n_tiles * n_images * chunks), threads, 0, hw, d_blocks, G, d_bits,
	            d_tile_off, d_total, with_heads, d_out, out_stride, chunks);
	return AKOD_OK;
}

__global__ void k_walk_blocks(const uint8_t* __restrict__ blob, uint64_t input_size, uint32_t n_tiles,
                              uint64_t* __restrict__ off, uint64_t* __restrict__ size)
{
	if (threadIdx.x != 0 || blockIdx.x != 0)
		return;
	uint64_t pos = 16;
	bool ok = true;
	for (uint32_t t = 0; t < n_tiles; t++)
	{
		uint64_t s = 0;
		if (ok && pos + 4 <= input_size)
		{
			s = (uint64_t)blob[pos] | ((uint64_t)blob[pos + 1] << 8) | ((uint64_t)blob[pos + 2] << 16) |
			    ((uint64_t)blob[pos + 3] << 24);
			if (s == 0 || pos + 4 + s > input_size)
				s = 0;
		}
		if (s == 0)
			ok = false;
		off[t] = pos + 4;
		size[t] = s;
		pos += 4 + s;
	}
}

extern "C" int akod_walk_blocks(akodContext* c, const uint8_t* d_blob, uint64_t input_size, uint32_t n_tiles,
                                uint64_t* d_off, uint64_t* d_size)
{
	akod_use(c);
	AKOD_LAUNCH(c, "walk_blocks", k_walk_blocks, 1, 32, 0, d_blob, input_size, n_tiles, d_off, d_size);
	return AKOD_OK;
}

// the same walk for n blobs at d_blobs + i*stride, one thread each; off/size are [n][tiles]
__global__ void k_walk_blocks_batch(const uint8_t* __restrict__ blobs, uint64_t stride, const uint64_t* __restrict__ input_size,
                                    uint32_t n_tiles, uint32_t n, uint64_t* __restrict__ off, uint64_t* __restrict__ size)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint8_t* blob = blobs + stride * i;
	const uint64_t in_size = input_size[i];
	uint64_t pos = 16;
	bool ok = true;
	for (uint32_t t = 0; t < n_tiles; t++)
	{
		uint64_t s = 0;
		if (ok && pos + 4 <= in_size)
		{
			s = (uint64_t)blob[pos] | ((uint64_t)blob[pos + 1] << 8) | ((uint64_t)blob[pos + 2] << 16) |
			    ((uint64_t)blob[pos + 3] << 24);
			if (s == 0 || pos + 4 + s > in_size)
				s = 0;
		}
		if (s == 0)
			ok = false;
		off[(uint64_t)i * n_tiles + t] = pos + 4;
		size[(uint64_t)i * n_tiles + t] = s;
		pos += 4 + s;
	}
}

extern "C" int akod_walk_blocks_batch(akodContext* c, const uint8_t* d_blobs, uint64_t stride, const uint64_t* d_input_size,
                                      uint32_t n_tiles, uint32_t n, uint64_t* d_off, uint64_t* d_size)
{
	akod_use(c);
	AKOD_LAUNCH(c, "walk_blocks", k_walk_blocks_batch, (n + 63) / 64, 64, 0, d_blobs, stride, d_input_size, n_tiles, n, d_off,
	            d_size);
	return AKOD_OK;
}
