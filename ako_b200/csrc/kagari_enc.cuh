// kagari_enc.cuh -- element-parallel Kagari (zig-zag + Elias gamma + RLE) encoder.
// Replaces akoKagariEncode / akoEliasEncodeStep / akoEliasEncodeEnd (reference library/kagari.c:59-116, :228-298).
//
// The reference walks the coefficient stream with a run counter. The same bits come out of a per-element rule
// (restated and pinned in oracle/ako_oracle.c:s_kagari_emit). For element i at 0-based position
// k inside its run of equal values, let c = 0 if k == 0 else ((k-1) mod 65534) + 1:
//     c <= 2          -> emit EV(a[i]) = gamma(zigzag16(a[i]) + 1)
//     c == 65534      -> emit gamma(65533), then c := 0           (run counter overflow, kagari.c:265-271)
//     last of its run and c >= 2 -> additionally emit gamma(c-1)   (kagari.c:275-279, :290-294)
// so every element contributes one bit string of at most 32 bits. Three passes over the stream:
//   1. k_kg_starts   : last run start per block            -> k_kg_scan_max (carry of run starts over blocks)
//   2. k_kg_lengths  : bits per block                      -> k_kg_scan_sum (64-bit bit offset of each block)
//   3. k_kg_pack     : codes OR-ed into a shared-memory bit buffer, written MSB-first as whole 32-bit words;
//                      the two words a block may share with its neighbours are merged with atomicOr
//                      (k_kg_zero_edges clears them first).
// Quantised streams are mostly long runs of zeros: a thread whose 8 values continue a run that also goes on
// after them emits nothing unless the run counter crosses 1, 2 or 65534 inside it, and skips all per-element work.
// Run starts are carried as (index + 1) in 32 bits (0 = none); the host refuses streams of 2^32 values or more.
// blockIdx.y is the image of a same-shape batch.
#pragma once

#include "common.cuh"

constexpr int KG_THREADS = 256;
constexpr int KG_ITEMS = 8;
constexpr int KG_BLOCK = KG_THREADS * KG_ITEMS; // 2048 values per CTA

struct KgCode
{
	uint32_t code;
	uint32_t len;
};

// gamma(v) for a uint16 v; v == 0 degenerates to a single 0 bit (kagari.c:38-45 with :61-62)
__device__ __forceinline__ KgCode kg_gamma(uint32_t v)
{
	KgCode r;
	if (v == 0)
	{
		r.code = 0;
		r.len = 1;
		return r;
	}
	const int b = 31 - __clz(v);
	r.code = v;
	r.len = 2 * b + 1;
	return r;
}

// EV(x): kagari.c:169-173 then +1 narrowed to uint16 by akoEliasEncodeStep's parameter (kagari.c:214-217)
__device__ __forceinline__ KgCode kg_value(int16_t x)
{
	const uint32_t zz = (uint32_t)(((int)x << 1) ^ ((int)x >> 15)) & 0xFFFFu;
	return kg_gamma((zz + 1) & 0xFFFFu);
}

// the whole bit string of one element; k = position in run, last = (next differs or end of stream)
__device__ __forceinline__ KgCode kg_element(int16_t a, uint32_t k, bool last)
{
	uint32_t c = (k == 0) ? 0u : ((k - 1) % 65534u) + 1u;
	KgCode out;
	out.code = 0;
	out.len = 0;
	if (c <= 2)
		out = kg_value(a);
	else if (c == 65534u)
	{
		out = kg_gamma(65533u);
		c = 0;
	}
	if (last && c >= 2)
	{
		const KgCode t = kg_gamma(c - 1);
		out.code = (out.len ? (out.code << t.len) : 0u) | t.code;
		out.len += t.len;
	}
	return out;
}

// what a thread knows about its KG_ITEMS values after loading them
struct KgChunk
{
	int16_t v[KG_ITEMS];
	int16_t after;
	int valid;           // how many of v[] exist
	bool has_after;      // an element follows the chunk
	uint32_t start_mask; // bit j: v[j] starts a run (differs from its predecessor, or is element 0)
	uint32_t last_start; // (index + 1) of the last run start inside the chunk, 0 if none
};

__device__ __forceinline__ KgChunk kg_load(const int16_t* __restrict__ in, uint64_t n, uint64_t base)
{
	KgChunk c;
	c.valid = 0;
	if (base + KG_ITEMS <= n)
	{
		// 128-bit load: base is a multiple of 8 and the stream is 16-byte aligned
		*reinterpret_cast<uint4*>(c.v) = __ldg(reinterpret_cast<const uint4*>(in + base));
		c.valid = KG_ITEMS;
	}
	else
	{
#pragma unroll
		for (int j = 0; j < KG_ITEMS; j++)
		{
			c.v[j] = 0;
			if (base + j < n)
			{
				c.v[j] = in[base + j];
				c.valid = j + 1;
			}
		}
	}
	const bool has_before = base > 0 && base <= n;
	const int16_t before = has_before ? __ldg(in + base - 1) : (int16_t)0;
	c.has_after = base + KG_ITEMS < n;
	c.after = c.has_after ? __ldg(in + base + KG_ITEMS) : (int16_t)0;
	c.start_mask = 0;
	c.last_start = 0;
#pragma unroll
	for (int j = 0; j < KG_ITEMS; j++)
	{
		const bool start = (j < c.valid) && ((j == 0) ? (!has_before || c.v[0] != before) : (c.v[j] != c.v[j - 1]));
		if (start)
		{
			c.start_mask |= 1u << j;
			c.last_start = (uint32_t)(base + j) + 1u;
		}
	}
	return c;
}

// pass 1: blk_start[b] = (largest i in block b that starts a run) + 1, or 0
__global__ void __launch_bounds__(KG_THREADS)
    k_kg_starts(const int16_t* __restrict__ in, uint64_t in_stride, uint64_t n, uint32_t* __restrict__ blk_start,
                uint32_t nblocks)
{
	__shared__ uint32_t sm[33];
	in += in_stride * blockIdx.y;
	blk_start += (uint64_t)nblocks * blockIdx.y;
	const uint64_t base = (uint64_t)blockIdx.x * KG_BLOCK + (uint64_t)threadIdx.x * KG_ITEMS;
	const KgChunk c = kg_load(in, n, base);
	uint32_t total;
	block_excl_max_u32(c.last_start, sm, &total);
	if (threadIdx.x == 0)
		blk_start[blockIdx.x] = total;
}

// exclusive max-scan over the blocks of one image (one CTA per image); identity 0
__global__ void __launch_bounds__(1024) k_kg_scan_max(uint32_t* __restrict__ blk, uint32_t nblocks)
{
	__shared__ uint32_t sm[33];
	blk += (uint64_t)nblocks * blockIdx.x;
	uint32_t carry = 0;
	for (uint32_t b0 = 0; b0 < nblocks; b0 += 1024)
	{
		const uint32_t b = b0 + threadIdx.x;
		const uint32_t v = (b < nblocks) ? blk[b] : 0;
		uint32_t total;
		const uint32_t ex = block_excl_max_u32(v, sm, &total);
		if (b < nblocks)
			blk[b] = max(carry, ex);
		carry = max(carry, total);
	}
}

// exclusive 64-bit sum-scan over the blocks of one image; total[img] receives the sum
__global__ void __launch_bounds__(1024)
    k_kg_scan_sum(const uint32_t* __restrict__ blk_bits, uint64_t* __restrict__ blk_off, uint32_t nblocks,
                  uint64_t* __restrict__ total)
{
	__shared__ uint32_t sm[33];
	blk_bits += (uint64_t)nblocks * blockIdx.x;
	blk_off += (uint64_t)nblocks * blockIdx.x;
	uint64_t carry = 0;
	for (uint32_t b0 = 0; b0 < nblocks; b0 += 1024)
	{
		const uint32_t b = b0 + threadIdx.x;
		const uint32_t v = (b < nblocks) ? blk_bits[b] : 0; // <= 65536 each, 1024 of them fit in 32 bits
		uint32_t tot;
		const uint32_t ex = block_excl_sum(v, sm, &tot);
		if (b < nblocks)
			blk_off[b] = carry + ex;
		carry += tot;
	}
	if (threadIdx.x == 0)
		total[blockIdx.x] = carry;
}

// shared by pass 2 and 3: the codes of a thread's KG_ITEMS elements; returns their total bit count.
// EMIT = false only counts bits.
template <bool EMIT>
__device__ __forceinline__ uint32_t kg_thread_codes(const KgChunk& c, uint64_t n, uint64_t base, uint32_t carry_start,
                                                    uint32_t* sm_max, KgCode codes[KG_ITEMS])
{
	// run start reaching into this thread = max(block carry, starts of earlier threads); all are (index + 1)
	uint32_t dummy;
	uint32_t run_start = max(carry_start, block_excl_max_u32(c.last_start, sm_max, &dummy));

	if (EMIT)
	{
#pragma unroll
		for (int j = 0; j < KG_ITEMS; j++)
		{
			codes[j].code = 0;
			codes[j].len = 0;
		}
	}

	// fast path: all 8 values continue one run that also continues after them
	if (c.valid == KG_ITEMS && c.start_mask == 0 && c.has_after && c.after == c.v[KG_ITEMS - 1])
	{
		const uint32_t k0 = (uint32_t)base + 1u - run_start; // position of v[0] in its run (>= 1)
		const uint32_t c0 = ((k0 - 1) % 65534u) + 1u;
		if (c0 >= 3 && c0 + (KG_ITEMS - 1) < 65534u)
			return 0; // the run counter stays strictly between 2 and 65534: nothing is emitted
	}

	uint32_t bits = 0;
#pragma unroll
	for (int j = 0; j < KG_ITEMS; j++)
	{
		if (j < c.valid)
		{
			if (c.start_mask & (1u << j))
				run_start = (uint32_t)(base + j) + 1u;
			const bool last = (j + 1 < c.valid) ? (c.v[j + 1] != c.v[j])
			                                    : ((base + j + 1 >= n) || (j + 1 == KG_ITEMS ? (c.after != c.v[j]) : true));
			const KgCode e = kg_element(c.v[j], (uint32_t)(base + j) + 1u - run_start, last);
			if (EMIT)
				codes[j] = e;
			bits += e.len;
		}
	}
	return bits;
}

// pass 2: bits emitted by each block
__global__ void __launch_bounds__(KG_THREADS)
    k_kg_lengths(const int16_t* __restrict__ in, uint64_t in_stride, uint64_t n, const uint32_t* __restrict__ blk_carry,
                 uint32_t* __restrict__ blk_bits, uint32_t nblocks)
{
	__shared__ uint32_t sm_max[33];
	__shared__ uint32_t sm_sum[33];
	in += in_stride * blockIdx.y;
	const uint64_t base = (uint64_t)blockIdx.x * KG_BLOCK + (uint64_t)threadIdx.x * KG_ITEMS;
	const KgChunk c = kg_load(in, n, base);
	const uint32_t bits =
	    kg_thread_codes<false>(c, n, base, blk_carry[(uint64_t)nblocks * blockIdx.y + blockIdx.x], sm_max, nullptr);
	uint32_t total;
	block_excl_sum(bits, sm_sum, &total);
	if (threadIdx.x == 0)
		blk_bits[(uint64_t)nblocks * blockIdx.y + blockIdx.x] = total;
}

// clears the (up to two) 32-bit words each block shares with its neighbours, and the final word
__global__ void __launch_bounds__(256)
    k_kg_zero_edges(const uint64_t* __restrict__ blk_off, const uint32_t* __restrict__ blk_bits, uint32_t nblocks,
                    uint8_t* __restrict__ out, uint64_t out_stride, uint64_t cap_bits)
{
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblocks)
		return;
	const uint64_t off = blk_off[(uint64_t)nblocks * blockIdx.y + b];
	const uint32_t bits = blk_bits[(uint64_t)nblocks * blockIdx.y + b];
	if (bits == 0 || off + bits > cap_bits)
		return;
	uint32_t* words = reinterpret_cast<uint32_t*>(out + out_stride * blockIdx.y);
	words[off >> 5] = 0;
	words[(off + bits - 1) >> 5] = 0;
}

// pass 3: pack
__global__ void __launch_bounds__(KG_THREADS)
    k_kg_pack(const int16_t* __restrict__ in, uint64_t in_stride, uint64_t n, const uint32_t* __restrict__ blk_carry,
              const uint64_t* __restrict__ blk_off, const uint32_t* __restrict__ blk_bits, uint32_t nblocks,
              uint8_t* __restrict__ out, uint64_t out_stride, uint64_t cap_bits)
{
	__shared__ uint32_t sm_max[33];
	__shared__ uint32_t sm_sum[33];
	__shared__ uint32_t bitbuf[KG_BLOCK + 2]; // 32 bits per value at most, +1 word of misalignment, +1 spill

	// a block that emits nothing (inside a long run) has nothing to do at all
	const uint32_t total = blk_bits[(uint64_t)nblocks * blockIdx.y + blockIdx.x];
	if (total == 0)
		return;
	const uint64_t g0 = blk_off[(uint64_t)nblocks * blockIdx.y + blockIdx.x];
	if (g0 + total > cap_bits) // would not fit: the caller reports the failure from the bit count
		return;

	in += in_stride * blockIdx.y;
	const uint64_t base = (uint64_t)blockIdx.x * KG_BLOCK + (uint64_t)threadIdx.x * KG_ITEMS;
	const KgChunk c = kg_load(in, n, base);
	KgCode codes[KG_ITEMS];
	const uint32_t bits =
	    kg_thread_codes<true>(c, n, base, blk_carry[(uint64_t)nblocks * blockIdx.y + blockIdx.x], sm_max, codes);

	for (int i = threadIdx.x; i < KG_BLOCK + 2; i += KG_THREADS)
		bitbuf[i] = 0;
	uint32_t check;
	const uint32_t excl = block_excl_sum(bits, sm_sum, &check); // has __syncthreads inside: bitbuf is clear after it

	if (bits)
	{
		uint32_t pos = (uint32_t)(g0 & 31) + excl; // bit position inside bitbuf
#pragma unroll
		for (int j = 0; j < KG_ITEMS; j++)
		{
			const uint32_t len = codes[j].len;
			if (len)
			{
				const uint32_t w = pos >> 5, sh = pos & 31;
				// MSB-first: bit 'pos' of the stream is bit (31 - pos%32) of word pos/32
				const uint64_t wide = (uint64_t)codes[j].code << (64 - sh - len);
				atomicOr(&bitbuf[w], (uint32_t)(wide >> 32));
				if (sh + len > 32)
					atomicOr(&bitbuf[w + 1], (uint32_t)wide);
				pos += len;
			}
		}
	}
	__syncthreads();

	uint32_t* words = reinterpret_cast<uint32_t*>(out + out_stride * blockIdx.y) + (g0 >> 5);
	const uint32_t nwords = (uint32_t)(((g0 & 31) + total + 31) >> 5);
	for (uint32_t i = threadIdx.x; i < nwords; i += KG_THREADS)
	{
		const uint32_t be = __byte_perm(bitbuf[i], 0, 0x0123); // bytes of the file are MSB-first
		if (i == 0 || i == nwords - 1)
			atomicOr(&words[i], be);
		else
			words[i] = be;
	}
}
