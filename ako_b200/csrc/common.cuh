// common.cuh -- context, launch accounting and block-level scan helpers shared by all kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "ako_device.h"

struct akodProfEntry
{
	const char* name;
	uint64_t launches;
	double ms;
	uint64_t bytes; // algorithmic bytes declared by the launchers (AKOD_BYTES), summed over launches
};

struct akodPending
{
	int entry;
	cudaEvent_t a, b;
};

struct akodContext
{
	int device;
	int sm_count;
	cudaStream_t stream;
	void* ws[AKOD_WS_COUNT];
	size_t ws_size[AKOD_WS_COUNT];
	void* mailbox; // pinned host, 64 KiB
	cudaEvent_t sync_event; // blocking-sync event (cudaEventBlockingSync), see akod_sync
	bool blocking_sync;     // wait for the stream by sleeping on sync_event instead of spinning in the driver
	// accounting
	bool profiling;
	uint64_t launch_count;
	uint64_t next_bytes; // algorithmic bytes of the next launch (consumed by AKOD_LAUNCH)
	bool small_attr_done; // cudaFuncSetAttribute for the small-pyramid kernels done on this device
	uint32_t strip4_attr_done; // ... for the fused level-0 kernels, one bit per wavelet
	std::vector<akodProfEntry> prof;
	std::vector<akodPending> pending;
	std::vector<cudaEvent_t> event_pool;
};

static inline int akod_cuda_status(cudaError_t e)
{
	if (e == cudaSuccess)
		return AKOD_OK;
	if (getenv("AKO_B200_DEBUG"))
		fprintf(stderr, "[ako_b200] CUDA error: %s\n", cudaGetErrorString(e));
	return (e == cudaErrorMemoryAllocation) ? AKOD_NOMEM : AKOD_ERROR;
}

#define AKOD_TRY(expr)                                   \
	do                                                   \
	{                                                    \
		const int akod_try_rc = akod_cuda_status(expr);  \
		if (akod_try_rc != AKOD_OK)                      \
			return akod_try_rc;                          \
	} while (0)

static inline int akod_prof_entry(akodContext* c, const char* name)
{
	for (size_t i = 0; i < c->prof.size(); i++)
		if (c->prof[i].name == name || strcmp(c->prof[i].name, name) == 0)
			return (int)i;
	c->prof.push_back(akodProfEntry{name, 0, 0.0, 0});
	return (int)c->prof.size() - 1;
}

static inline cudaEvent_t akod_event_get(akodContext* c)
{
	if (!c->event_pool.empty())
	{
		cudaEvent_t e = c->event_pool.back();
		c->event_pool.pop_back();
		return e;
	}
	cudaEvent_t e;
	cudaEventCreate(&e);
	return e;
}

// Declares the algorithmic bytes of the launch that follows; roofline accounting only.
#define AKOD_BYTES(ctx, n) ((ctx)->next_bytes = (uint64_t)(n))

// Every kernel of the library is launched through this macro: it counts the launch (always) and,
// when profiling is enabled, brackets it with CUDA events on the context's stream.
#define AKOD_LAUNCH(ctx, label, kernel, grid, block, smem, ...)                              \
	do                                                                                       \
	{                                                                                        \
		akodContext* akod_l_c = (ctx);                                                       \
		const int akod_l_e = akod_prof_entry(akod_l_c, label);                               \
		akod_l_c->prof[akod_l_e].launches++;                                                 \
		akod_l_c->prof[akod_l_e].bytes += akod_l_c->next_bytes;                              \
		akod_l_c->next_bytes = 0;                                                            \
		akod_l_c->launch_count++;                                                            \
		akodPending akod_l_p;                                                                \
		if (akod_l_c->profiling)                                                             \
		{                                                                                    \
			akod_l_p.entry = akod_l_e;                                                       \
			akod_l_p.a = akod_event_get(akod_l_c);                                           \
			akod_l_p.b = akod_event_get(akod_l_c);                                           \
			cudaEventRecord(akod_l_p.a, akod_l_c->stream);                                   \
		}                                                                                    \
		kernel<<<(grid), (block), (smem), akod_l_c->stream>>>(__VA_ARGS__);                  \
		if (akod_l_c->profiling)                                                             \
		{                                                                                    \
			cudaEventRecord(akod_l_p.b, akod_l_c->stream);                                   \
			akod_l_c->pending.push_back(akod_l_p);                                           \
		}                                                                                    \
		AKOD_TRY(cudaGetLastError());                                                        \
	} while (0)

// ---------------------------------------------------------------- device helpers

#define AKOD_FULL_MASK 0xffffffffu

template <typename T>
__device__ __forceinline__ T akod_max(T a, T b)
{
	return a > b ? a : b;
}

// Inclusive warp scans
__device__ __forceinline__ uint32_t warp_incl_sum(uint32_t v)
{
	const int lane = threadIdx.x & 31;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const uint32_t o = __shfl_up_sync(AKOD_FULL_MASK, v, d);
		if (lane >= d)
			v += o;
	}
	return v;
}

__device__ __forceinline__ long long warp_incl_max(long long v)
{
	const int lane = threadIdx.x & 31;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const long long o = __shfl_up_sync(AKOD_FULL_MASK, v, d);
		if (lane >= d)
			v = akod_max(v, o);
	}
	return v;
}

// Block-wide exclusive sum of one uint32 per thread; returns the exclusive prefix, *total gets the block sum.
// sm must hold 33 uint32. blockDim.x must be a multiple of 32, <= 1024.
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t* sm, uint32_t* total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	const uint32_t incl = warp_incl_sum(v);
	if (lane == 31)
		sm[wid] = incl;
	__syncthreads();
	if (wid == 0)
	{
		uint32_t w = (lane < nw) ? sm[lane] : 0;
		const uint32_t wi = warp_incl_sum(w);
		sm[lane] = wi - w;
		if (lane == 31)
			sm[32] = wi;
	}
	__syncthreads();
	const uint32_t r = sm[wid] + incl - v;
	*total = sm[32];
	__syncthreads();
	return r;
}

// Single-use variants for kernels that scan once per shared array: ONE barrier each. Every warp folds the NW warp
// totals itself (one load per lane + two REDUX) instead of waiting for warp 0 to scan them, and there is no trailing barrier:
// the caller must not write sm again before a barrier of its own.
template <int NW>
__device__ __forceinline__ uint32_t block_excl_sum_once(uint32_t v, uint32_t* sm, uint32_t* total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t incl = warp_incl_sum(v);
	if (lane == 31)
		sm[wid] = incl;
	__syncthreads();
	// every warp folds the NW warp totals itself: one shared load per lane and two warp reductions (REDUX)
	const uint32_t x = (lane < NW) ? sm[lane] : 0u;
	const uint32_t all = __reduce_add_sync(AKOD_FULL_MASK, x);
	const uint32_t before = __reduce_add_sync(AKOD_FULL_MASK, (lane < wid) ? x : 0u);
	*total = all;
	return before + incl - v;
}

// Block-wide exclusive max of one int64 per thread (identity = -1); *total gets the block max.
// sm must hold 33 long long.
__device__ __forceinline__ long long block_excl_max(long long v, long long* sm, long long* total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	const long long incl = warp_incl_max(v);
	long long excl = __shfl_up_sync(AKOD_FULL_MASK, incl, 1);
	if (lane == 0)
		excl = -1;
	if (lane == 31)
		sm[wid] = incl;
	__syncthreads();
	if (wid == 0)
	{
		long long w = (lane < nw) ? sm[lane] : -1;
		const long long wi = warp_incl_max(w);
		long long we = __shfl_up_sync(AKOD_FULL_MASK, wi, 1);
		if (lane == 0)
			we = -1;
		sm[lane] = we;
		if (lane == 31)
			sm[32] = wi;
	}
	__syncthreads();
	const long long r = akod_max(sm[wid], excl);
	*total = sm[32];
	__syncthreads();
	return r;
}

// Block-wide exclusive max of one uint32 per thread (identity 0); *total gets the block max. sm: 33 uint32.
__device__ __forceinline__ uint32_t block_excl_max_u32(uint32_t v, uint32_t* sm, uint32_t* total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const uint32_t o = __shfl_up_sync(AKOD_FULL_MASK, incl, d);
		if (lane >= d)
			incl = max(incl, o);
	}
	uint32_t excl = __shfl_up_sync(AKOD_FULL_MASK, incl, 1);
	if (lane == 0)
		excl = 0;
	if (lane == 31)
		sm[wid] = incl;
	__syncthreads();
	if (wid == 0)
	{
		uint32_t w = (lane < nw) ? sm[lane] : 0;
		uint32_t wi = w;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const uint32_t o = __shfl_up_sync(AKOD_FULL_MASK, wi, d);
			if (lane >= d)
				wi = max(wi, o);
		}
		uint32_t we = __shfl_up_sync(AKOD_FULL_MASK, wi, 1);
		if (lane == 0)
			we = 0;
		sm[lane] = we;
		if (lane == 31)
			sm[32] = wi;
	}
	__syncthreads();
	const uint32_t r = max(sm[wid], excl);
	*total = sm[32];
	__syncthreads();
	return r;
}

// Block-wide "last nonzero value of the threads before me" for values that are non-decreasing in thread order
// wherever they are nonzero (run starts carried as index + 1), which makes it an exclusive max: one ballot and one
// shuffle per warp instead of a five-step scan. *total = last nonzero value of the block. sm: 33 uint32.
__device__ __forceinline__ uint32_t block_excl_last_start(uint32_t v, uint32_t* sm, uint32_t* total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	const uint32_t mask = __ballot_sync(AKOD_FULL_MASK, v != 0);
	const uint32_t lower = mask & ((1u << lane) - 1u);
	const uint32_t from_lower = __shfl_sync(AKOD_FULL_MASK, v, lower ? 31 - __clz(lower) : 0);
	const uint32_t warp_last = __shfl_sync(AKOD_FULL_MASK, v, mask ? 31 - __clz(mask) : 0);
	if (lane == 0)
		sm[wid] = mask ? warp_last : 0u;
	__syncthreads();
	uint32_t carry = 0, all = 0;
	for (int w = 0; w < nw; w++)
	{
		const uint32_t x = sm[w];
		all = max(all, x);
		if (w < wid)
			carry = max(carry, x);
	}
	*total = all;
	__syncthreads();
	return lower ? from_lower : carry;
}

template <int NW>
__device__ __forceinline__ uint32_t block_excl_last_start_once(uint32_t v, uint32_t* sm, uint32_t* total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t mask = __ballot_sync(AKOD_FULL_MASK, v != 0);
	const uint32_t lower = mask & ((1u << lane) - 1u);
	const uint32_t from_lower = __shfl_sync(AKOD_FULL_MASK, v, lower ? 31 - __clz(lower) : 0);
	const uint32_t warp_last = __shfl_sync(AKOD_FULL_MASK, v, mask ? 31 - __clz(mask) : 0);
	if (lane == 0)
		sm[wid] = mask ? warp_last : 0u;
	__syncthreads();
	const uint32_t x = (lane < NW) ? sm[lane] : 0u;
	const uint32_t all = __reduce_max_sync(AKOD_FULL_MASK, x);
	const uint32_t carry = __reduce_max_sync(AKOD_FULL_MASK, (lane < wid) ? x : 0u);
	*total = all;
	return lower ? from_lower : carry;
}
