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
	uint32_t c = k; // run counter: 0 for the first element of a run, then 1..65534 cyclically
	if (k > 65534u)
		c = ((k - 1) % 65534u) + 1u;
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

// Must be called by whole warps (neighbouring values travel by shuffle): thread t of a warp holds the 8 values
// that follow those of thread t-1.
__device__ __forceinline__ KgChunk kg_load(const int16_t* __restrict__ in, uint64_t n, uint64_t base)
{
	KgChunk c;
	const int lane = threadIdx.x & 31;
	uint32_t w[4] = {0, 0, 0, 0};
	const bool full = base + KG_ITEMS <= n;
	if (full)
	{
		// 128-bit load: base is a multiple of 8 and the stream is 16-byte aligned
		const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + base));
		w[0] = q.x;
		w[1] = q.y;
		w[2] = q.z;
		w[3] = q.w;
	}
	// the value before / after the chunk comes from the neighbouring lane when that lane holds a full chunk
	uint32_t before = __shfl_up_sync(AKOD_FULL_MASK, w[3], 1) >> 16;
	uint32_t after = __shfl_down_sync(AKOD_FULL_MASK, w[0], 1) & 0xFFFFu;
	const bool next_full = base + 2 * KG_ITEMS <= n;
	const bool has_before = base > 0 && base <= n;
	c.has_after = base + KG_ITEMS < n;
	if (lane == 0 || !full)
		before = has_before ? (uint32_t)(uint16_t)__ldg(in + base - 1) : 0u;
	if (lane == 31 || !next_full)
		after = c.has_after ? (uint32_t)(uint16_t)__ldg(in + base + KG_ITEMS) : 0u;
	c.after = (int16_t)after;

	if (full)
	{
		*reinterpret_cast<uint4*>(c.v) = make_uint4(w[0], w[1], w[2], w[3]);
		c.valid = KG_ITEMS;
		// element j differs from its predecessor <=> half j of (w ^ w shifted by one element) is nonzero
		uint32_t mask = 0;
		uint32_t prev_word = before << 16;
#pragma unroll
		for (int i = 0; i < 4; i++)
		{
			const uint32_t d = w[i] ^ __funnelshift_l(prev_word, w[i], 16);
			mask |= ((d & 0xFFFFu) ? 1u : 0u) << (2 * i);
			mask |= ((d >> 16) ? 1u : 0u) << (2 * i + 1);
			prev_word = w[i];
		}
		if (!has_before)
			mask |= 1u; // element 0 of the stream starts a run
		c.start_mask = mask;
		c.last_start = mask ? (uint32_t)base + (31 - __clz(mask)) + 1u : 0u;
		return c;
	}

	c.valid = 0;
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
	c.start_mask = 0;
	c.last_start = 0;
#pragma unroll
	for (int j = 0; j < KG_ITEMS; j++)
	{
		const bool start =
		    (j < c.valid) && ((j == 0) ? (!has_before || c.v[0] != (int16_t)before) : (c.v[j] != c.v[j - 1]));
		if (start)
		{
			c.start_mask |= 1u << j;
			c.last_start = (uint32_t)(base + j) + 1u;
		}
	}
	return c;
}

// pass 1: blk_own[b] = (largest i in block b that starts a run) + 1, or 0; blk_first[b] = 1 when the block's first
// value starts a run. A CTA takes KG_STARTS_PER_CTA consecutive blocks. Full blocks take a lean path (one 128-bit
// load per block, all issued before any arithmetic; four XORs against the stream shifted by one element; a warp
// REDUX; one barrier for all blocks); the last, partial block of a stream goes through kg_load.
constexpr int KG_STARTS_PER_CTA = 2;

// last run start of this thread's 8 values at 'base' as (index + 1), 0 if none; *first_differs: value 0 starts a run
__device__ __forceinline__ uint32_t kg_lean_last_start(const int16_t* __restrict__ in, uint64_t base, uint4 q, bool* first_differs)
{
	const int lane = threadIdx.x & 31;
	// word whose HIGH half is the element before this thread's first one
	uint32_t prev = __shfl_up_sync(AKOD_FULL_MASK, q.w, 1);
	if (lane == 0)
		prev = (base > 0) ? ((uint32_t)(uint16_t)__ldg(in + base - 1) << 16) : (~q.x << 16); // element 0 starts a run
	// half j of d is nonzero <=> that element differs from its predecessor
	const uint32_t d0 = q.x ^ __funnelshift_l(prev, q.x, 16);
	const uint32_t d1 = q.y ^ __funnelshift_l(q.x, q.y, 16);
	const uint32_t d2 = q.z ^ __funnelshift_l(q.y, q.z, 16);
	const uint32_t d3 = q.w ^ __funnelshift_l(q.z, q.w, 16);
	uint32_t last = 0;
	if (d3)
		last = (d3 >> 16) ? 8 : 7;
	else if (d2)
		last = (d2 >> 16) ? 6 : 5;
	else if (d1)
		last = (d1 >> 16) ? 4 : 3;
	else if (d0)
		last = (d0 >> 16) ? 2 : 1;
	*first_differs = (d0 & 0xFFFFu) != 0;
	return last ? (uint32_t)base + last : 0u;
}

__global__ void __launch_bounds__(KG_THREADS)
    k_kg_starts(const int16_t* __restrict__ in, uint64_t in_stride, uint64_t n, uint32_t* __restrict__ blk_own,
                uint8_t* __restrict__ blk_first, uint32_t nblocks)
{
	__shared__ uint32_t sm[KG_STARTS_PER_CTA][33];
	in += in_stride * blockIdx.y;
	blk_own += (uint64_t)nblocks * blockIdx.y;
	blk_first += (uint64_t)nblocks * blockIdx.y;
	const uint32_t b0 = blockIdx.x * KG_STARTS_PER_CTA;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

	if ((uint64_t)(b0 + KG_STARTS_PER_CTA) * KG_BLOCK <= n)
	{
		uint4 q[KG_STARTS_PER_CTA];
#pragma unroll
		for (int k = 0; k < KG_STARTS_PER_CTA; k++)
			q[k] = __ldg(reinterpret_cast<const uint4*>(in + (uint64_t)(b0 + k) * KG_BLOCK + (uint64_t)threadIdx.x * KG_ITEMS));
		bool first[KG_STARTS_PER_CTA];
#pragma unroll
		for (int k = 0; k < KG_STARTS_PER_CTA; k++)
		{
			const uint64_t base = (uint64_t)(b0 + k) * KG_BLOCK + (uint64_t)threadIdx.x * KG_ITEMS;
			const uint32_t mine = kg_lean_last_start(in, base, q[k], &first[k]);
			const uint32_t warp_max = __reduce_max_sync(AKOD_FULL_MASK, mine);
			if (lane == 0)
				sm[k][wid] = warp_max;
		}
		__syncthreads();
		if (threadIdx.x == 0)
		{
#pragma unroll
			for (int k = 0; k < KG_STARTS_PER_CTA; k++)
			{
				uint32_t total = 0;
#pragma unroll
				for (int w = 0; w < KG_THREADS / 32; w++)
					total = max(total, sm[k][w]);
				blk_own[b0 + k] = total;
				blk_first[b0 + k] = (uint8_t)first[k];
			}
		}
		return;
	}

	// the end of the stream: block by block through the general loader
	for (int k = 0; k < KG_STARTS_PER_CTA; k++)
	{
		const uint32_t b = b0 + k;
		if (b >= nblocks)
			break;
		const uint64_t base = (uint64_t)b * KG_BLOCK + (uint64_t)threadIdx.x * KG_ITEMS;
		const KgChunk c = kg_load(in, n, base);
		uint32_t total;
		block_excl_last_start_once<KG_THREADS / 32>(c.last_start, sm[k], &total);
		if (threadIdx.x == 0)
		{
			blk_own[b] = total;
			blk_first[b] = (uint8_t)(c.start_mask & 1u);
		}
	}
}

// The two scans over the blocks of an image run in one CTA per image; a thread takes KG_SCAN_ITEMS consecutive
// blocks per round, so that a 16384 x 16384 image (half a million blocks) is 32 rounds of one block-wide scan each
// instead of 512.
constexpr int KG_SCAN_ITEMS = 16;

// exclusive max-scan over the blocks of one image (one CTA per image); identity 0
__global__ void __launch_bounds__(1024)
    k_kg_scan_max(const uint32_t* __restrict__ blk, uint32_t* __restrict__ blk_carry, uint32_t nblocks)
{
	__shared__ uint32_t sm[33];
	blk += (uint64_t)nblocks * blockIdx.x;
	blk_carry += (uint64_t)nblocks * blockIdx.x;
	uint32_t carry = 0;
	for (uint32_t b0 = 0; b0 < nblocks; b0 += 1024 * KG_SCAN_ITEMS)
	{
		const uint32_t first = b0 + threadIdx.x * KG_SCAN_ITEMS;
		uint32_t v[KG_SCAN_ITEMS];
		uint32_t last = 0; // last nonzero of this thread's blocks (run starts are increasing: "last nonzero" = max)
#pragma unroll
		for (int i = 0; i < KG_SCAN_ITEMS; i++)
		{
			v[i] = (first + i < nblocks) ? blk[first + i] : 0;
			last = v[i] ? v[i] : last;
		}
		uint32_t total;
		uint32_t run = max(carry, block_excl_last_start(last, sm, &total));
#pragma unroll
		for (int i = 0; i < KG_SCAN_ITEMS; i++)
		{
			if (first + i < nblocks)
				blk_carry[first + i] = run;
			run = v[i] ? v[i] : run;
		}
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
	for (uint32_t b0 = 0; b0 < nblocks; b0 += 1024 * KG_SCAN_ITEMS)
	{
		const uint32_t first = b0 + threadIdx.x * KG_SCAN_ITEMS;
		uint32_t v[KG_SCAN_ITEMS];
		uint32_t mine = 0; // <= 65536 each, 16384 of them fit in 32 bits
#pragma unroll
		for (int i = 0; i < KG_SCAN_ITEMS; i++)
		{
			v[i] = (first + i < nblocks) ? blk_bits[first + i] : 0;
			mine += v[i];
		}
		uint32_t tot;
		uint64_t at = carry + block_excl_sum(mine, sm, &tot);
#pragma unroll
		for (int i = 0; i < KG_SCAN_ITEMS; i++)
		{
			if (first + i < nblocks)
				blk_off[first + i] = at;
			at += v[i];
		}
		carry += tot;
	}
	if (threadIdx.x == 0)
		total[blockIdx.x] = carry;
}

// shared by pass 2 and 3: the codes of a thread's KG_ITEMS elements; returns their total bit count.
// EMIT = false only counts bits.
template <bool EMIT>
__device__ __forceinline__ uint32_t kg_codes(const KgChunk& c, uint64_t n, uint64_t base, uint32_t run_start,
                                             KgCode codes[KG_ITEMS]);

template <bool EMIT>
__device__ __forceinline__ uint32_t kg_thread_codes(const KgChunk& c, uint64_t n, uint64_t base, uint32_t carry_start,
                                                    uint32_t* sm_max, KgCode codes[KG_ITEMS])
{
	// run start reaching into this thread = max(block carry, starts of earlier threads); all are (index + 1)
	uint32_t dummy;
	const uint32_t run_start = max(carry_start, block_excl_last_start_once<KG_THREADS / 32>(c.last_start, sm_max, &dummy));
	return kg_codes<EMIT>(c, n, base, run_start, codes);
}

// run_start = (index + 1) of the start of the run that reaches into this thread's first value
template <bool EMIT>
__device__ __forceinline__ uint32_t kg_codes(const KgChunk& c, uint64_t n, uint64_t base, uint32_t run_start,
                                             KgCode codes[KG_ITEMS])
{

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
		uint32_t c0 = k0;
		if (k0 > 65534u)
			c0 = ((k0 - 1) % 65534u) + 1u;
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

// Every block owns a slot of KG_SLOT_WORDS words in a scratch buffer: pass 2 leaves the block's bit string there
// (packed from bit 0) when it fits, and pass 3 only has to shift it into place. Blocks whose string is longer
// (more than 8 bits per value on average) are re-encoded by pass 3 as before.
constexpr uint32_t KG_SLOT_WORDS = KG_BLOCK / 4;
constexpr uint32_t KG_SLOT_BITS = KG_SLOT_WORDS * 32;

// ORs the codes of one thread into a shared-memory bit buffer; 'pos' is the bit position of the first code
__device__ __forceinline__ void kg_put_codes(uint32_t* bitbuf, uint32_t pos, const KgCode codes[KG_ITEMS])
{
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

// ------------------------------------------------------------------------------------------------
// Pass 2 works from bit masks instead of walking the elements. For a thread's 8 values, S = start_mask (bit j: value
// j starts a run) and c0 = the run counter of value 0 when it continues a run. With no counter overflow inside the
// thread (c0 + 7 < 65534, all but one thread in 8192 of a long run):
//     value j emits EV(a[j])            <=>  its counter is <= 2  <=>  a start at j, j-1 or j-2 (or c0 + j <= 2)
//     value j emits its run's length    <=>  j is the last of its run (a start at j+1, or the end of the stream)
//                                            and its counter is >= 2 (no start at j nor at j-1)
// so V and R below name the emitters and only those are looked at: a quantised stream has one emitter per ~16
// values. The other threads (a counter overflow in reach, or the partial chunk at the end of the stream) take the
// element-by-element rule of kg_element.
struct KgMasks
{
	uint32_t V, R; // bit j: value j emits its value / the length of its run
	uint32_t c0;   // run counter of value 0 when it continues a run
	bool slow;     // take the element-by-element path
};

__device__ __forceinline__ KgMasks kg_masks(const KgChunk& c, uint64_t base, uint32_t run_start)
{
	KgMasks m;
	const uint32_t S = c.start_mask;
	const bool cont = (S & 1u) == 0; // value 0 continues the run that starts at run_start - 1
	uint32_t c0 = 0;
	if (cont)
	{
		const uint32_t k0 = (uint32_t)base + 1u - run_start; // >= 1
		c0 = (k0 > 65534u) ? ((k0 - 1u) % 65534u) + 1u : k0;
	}
	m.c0 = c0;
	m.slow = c.valid != KG_ITEMS || (cont && c0 + (KG_ITEMS - 1) >= 65534u);
	uint32_t V = (S | (S << 1) | (S << 2)) & 0xFFu;
	if (cont)
		V |= (c0 == 1u) ? 3u : (c0 == 2u) ? 1u : 0u;
	const uint32_t last = (!c.has_after || c.after != c.v[KG_ITEMS - 1]) ? 0x80u : 0u; // value 7 ends its run
	uint32_t K2 = ~(S | (S << 1)) & 0xFFu;                                               // counter >= 2
	if (cont && c0 < 2u)
		K2 &= ~1u;
	m.V = V;
	m.R = ((S >> 1) | last) & K2;
	return m;
}

// a bit string under construction in registers: codes are appended at the low end of a 64-bit accumulator and
// whole 32-bit words leave for the shared buffer as they fill up. The first and the last word of a thread's string
// are shared with its neighbours (atomicOr); the words in between are its own.
struct KgSink
{
	uint32_t* buf;
	uint64_t acc;
	uint32_t n, w; // bits pending in acc (< 32 between calls), next word of buf
	bool first;
	__device__ __forceinline__ void open(uint32_t* bitbuf, uint32_t pos)
	{
		buf = bitbuf;
		acc = 0;
		n = pos & 31u; // the bits before 'pos' in its word belong to the threads before: zeros here
		w = pos >> 5;
		first = true;
	}
	__device__ __forceinline__ void put(uint32_t code, uint32_t len)
	{
		acc = (acc << len) | code;
		n += len;
		if (n >= 32u)
		{
			n -= 32u;
			const uint32_t word = (uint32_t)(acc >> n);
			if (first)
				atomicOr(&buf[w], word);
			else
				buf[w] = word;
			first = false;
			w++;
		}
	}
	__device__ __forceinline__ void close()
	{
		if (n)
			atomicOr(&buf[w], (uint32_t)(acc << (32u - n)));
	}
};

struct KgCount
{
	uint32_t bits;
	__device__ __forceinline__ void put(uint32_t, uint32_t len) { bits += len; }
};

// The emitters of a chunk as a list of Elias-gamma arguments ("payloads", 16 bits each: zigzag(value) + 1 for a value,
// counter - 1 for a run length; 0 stands for the degenerate one-bit code of -32768, kagari.c:38-45 with :214-217), in
// stream order. PUSH(payload) is called once per code.
template <typename PUSH>
__device__ __forceinline__ void kg_emitters(const KgChunk& c, const KgMasks& m, uint64_t n, uint64_t base, uint32_t run_start,
                                            PUSH push)
{
	// the lane's 8 values as two 64-bit words: value j is 16 bits at 16 * (j & 3) of word j >> 2
	const uint4 q = *reinterpret_cast<const uint4*>(c.v);
	const uint64_t lo64 = ((uint64_t)q.y << 32) | q.x, hi64 = ((uint64_t)q.w << 32) | q.z;
	auto value = [&](int j) { return (int16_t)(uint16_t)(((j & 4) ? hi64 : lo64) >> (16 * (j & 3))); };
	if (m.slow)
	{
		// element by element (kg_element's rule): a counter overflow in reach, or the partial chunk at the end
#pragma unroll 1
		for (int j = 0; j < c.valid; j++)
		{
			if (c.start_mask & (1u << j))
				run_start = (uint32_t)(base + j) + 1u;
			const int16_t a = value(j);
			const bool last = (j + 1 < c.valid) ? (value(j + 1) != a)
			                                    : ((base + j + 1 >= n) || (j + 1 == KG_ITEMS ? (c.after != a) : true));
			const uint32_t k = (uint32_t)(base + j) + 1u - run_start;
			uint32_t cnt = k;
			if (k > 65534u)
				cnt = ((k - 1) % 65534u) + 1u;
			if (cnt <= 2)
				push(((uint32_t)(((int)a << 1) ^ ((int)a >> 15)) + 1u) & 0xFFFFu);
			else if (cnt == 65534u)
			{
				push(65533u);
				cnt = 0;
			}
			if (last && cnt >= 2)
				push(cnt - 1u);
		}
		return;
	}
	uint32_t U = m.V | m.R;
#pragma unroll 1
	while (U)
	{
		const int j = __ffs(U) - 1;
		U &= U - 1;
		if (m.V & (1u << j))
		{
			const int a = value(j);
			push(((uint32_t)((a << 1) ^ (a >> 15)) + 1u) & 0xFFFFu);
		}
		if (m.R & (1u << j))
		{
			// run counter of value j: distance to the last start at or before j, else c0 + j
			const uint32_t before = c.start_mask & ((2u << j) - 1u);
			const uint32_t cj = before ? (uint32_t)j - (31u - (uint32_t)__clz(before)) : m.c0 + (uint32_t)j;
			push(cj - 1u); // cj >= 2
		}
	}
}

__device__ __forceinline__ uint32_t kg_payload_len(uint32_t p)
{
	return 63u - 2u * (uint32_t)__clz(p | 1u); // 2 * floor(log2 p) + 1; the degenerate p == 0 is one bit
}

// pass 2: bits emitted by each block (+ the bit string itself into the block's slot when it fits).
// A WARP owns a block (and KGL_BLOCKS_PER_WARP consecutive blocks, one after the other): it walks the block's 2048
// values in eight steps of 256, carrying the run start and the bit position from step to step in registers, so the
// pass has no __syncthreads and no shared-memory scans at all -- only ballots and shuffles. With one CTA per block the
// pass was bound by the latency chain of a short-lived CTA (summary words -> values -> three barriers) and by the
// 300 000 CTAs per step that only find out that their block lies inside a run. The values of step i+1 are fetched
// before step i is worked on. Every lane assembles the bit string of its own eight values in a 64-bit register and ORs
// it into the warp's bit buffer (see the step body).
constexpr int KGL_WARPS = KG_THREADS / 32;
constexpr int KGL_BLOCKS_PER_WARP = 4; // at most: the launcher takes fewer when the batch is too small to fill the GPU so
constexpr int KGL_STEPS = KG_BLOCK / (32 * KG_ITEMS); // 8

// the raw loads of one step: the lane's 8 values and, for the lanes at the warp's edges, the neighbouring values
struct KgFetch
{
	uint4 q;
	uint32_t before, after; // lane 0 / lane 31 only
};

__device__ __forceinline__ KgFetch kg_fetch(const int16_t* __restrict__ in, uint64_t n, uint64_t base, int lane)
{
	KgFetch f;
	f.q = make_uint4(0, 0, 0, 0);
	f.before = f.after = 0;
	if (base + KG_ITEMS <= n)
		f.q = __ldg(reinterpret_cast<const uint4*>(in + base));
	if (lane == 0 && base > 0 && base <= n)
		f.before = (uint32_t)(uint16_t)__ldg(in + base - 1);
	if (lane == 31 && base + KG_ITEMS < n)
		f.after = (uint32_t)(uint16_t)__ldg(in + base + KG_ITEMS);
	return f;
}

// kg_load from prefetched registers (full chunks only: the caller sends the end of the stream through kg_load)
__device__ __forceinline__ KgChunk kg_chunk_from(const KgFetch& f, uint64_t n, uint64_t base, int lane)
{
	KgChunk c;
	const uint32_t w[4] = {f.q.x, f.q.y, f.q.z, f.q.w};
	uint32_t before = __shfl_up_sync(AKOD_FULL_MASK, w[3], 1) >> 16;
	uint32_t after = __shfl_down_sync(AKOD_FULL_MASK, w[0], 1) & 0xFFFFu;
	if (lane == 0)
		before = f.before;
	if (lane == 31)
		after = f.after;
	c.has_after = base + KG_ITEMS < n;
	c.after = (int16_t)after;
	*reinterpret_cast<uint4*>(c.v) = f.q;
	c.valid = KG_ITEMS;
	uint32_t mask = 0;
	uint32_t prev_word = before << 16;
#pragma unroll
	for (int i = 0; i < 4; i++)
	{
		// halves of d: value 2i against 2i-1, value 2i+1 against 2i; min(d, 1) per half (VIMNMX.U16x2), then both bits
		const uint32_t d = w[i] ^ __funnelshift_l(prev_word, w[i], 16);
		const uint32_t ne = __vminu2(d, 0x00010001u);
		mask |= ((ne | (ne >> 15)) & 3u) << (2 * i);
		prev_word = w[i];
	}
	if (base == 0)
		mask |= 1u; // element 0 of the stream starts a run
	c.start_mask = mask;
	c.last_start = mask ? (uint32_t)base + (31 - __clz(mask)) + 1u : 0u;
	return c;
}

// The codes of a lane's own emitters in stream order, straight from its registers (value j first, then the length of
// the run that ends at j): the general sink of a lane whose bit string does not fit 64 bits. PUT(payload, length).
template <typename PUT>
__device__ __forceinline__ void kg_lane_codes(const KgChunk& c, const KgMasks& m, PUT put)
{
	const uint4 q = *reinterpret_cast<const uint4*>(c.v);
	const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
	for (int j = 0; j < KG_ITEMS; j++)
	{
		if (m.V & (1u << j))
		{
			const int a = (j & 1) ? ((int)w[j >> 1] >> 16) : (int)(short)w[j >> 1];
			const uint32_t p = ((uint32_t)((a << 1) ^ (a >> 15)) + 1u) & 0xFFFFu;
			put(p, kg_payload_len(p));
		}
		if (m.R & (1u << j))
		{
			const uint32_t before = c.start_mask & ((2u << j) - 1u);
			const uint32_t cj = before ? (uint32_t)j - (31u - (uint32_t)__clz(before)) : m.c0 + (uint32_t)j;
			put(cj - 1u, kg_payload_len(cj - 1u)); // cj >= 2
		}
	}
}

// one block by one warp; returns the block's bit count. WHOLE = false: the last, partial block of a stream, whose
// chunks come through the general loader (kept out of line: it is one block per stream and would only cost the
// common path registers)
template <bool WHOLE>
__device__ __forceinline__ uint32_t kgl_block_body(const int16_t* __restrict__ in, uint64_t n, uint64_t first, uint32_t carry_start,
                                                   uint32_t* bitbuf, KgFetch f, int lane)
{
	uint32_t run_carry = carry_start; // (index + 1) of the last run start before the current step
	uint32_t bitpos = 0;
	bool overflow = false;
#pragma unroll 1
	for (int step = 0; step < KGL_STEPS; step++)
	{
		const uint64_t base = first + (uint64_t)step * (32 * KG_ITEMS) + (uint64_t)lane * KG_ITEMS;
		const KgFetch cur = f;
		if (WHOLE && step + 1 < KGL_STEPS)
			f = kg_fetch(in, n, base + 32 * KG_ITEMS, lane);
		const KgChunk c = WHOLE ? kg_chunk_from(cur, n, base, lane) : kg_load(in, n, base);
		// run start reaching into this lane: the last start of the lanes before it, else the carry
		const uint32_t any = __ballot_sync(AKOD_FULL_MASK, c.last_start != 0);
		const uint32_t lower = any & ((1u << lane) - 1u);
		const uint32_t from_lower = __shfl_sync(AKOD_FULL_MASK, c.last_start, lower ? 31 - __clz(lower) : 0);
		const uint32_t warp_last = __shfl_sync(AKOD_FULL_MASK, c.last_start, any ? 31 - __clz(any) : 0);
		const uint32_t run_start = lower ? from_lower : run_carry;
		if (any)
			run_carry = warp_last;

		// ---- the step's emitters. Every lane codes its own: its whole bit string is assembled in a 64-bit register in
		// ONE unrolled, branch-free pass over its eight values (a run length follows its value only where some lane of
		// the warp has one at that position), and after the warp scan of the lengths it is placed with at most three
		// atomicOr. (Round 1 compacted the emitters of a step into a list and dealt them over the lanes again: balanced,
		// but twice through shared memory, with divergent loops on both sides -- 2.5 x the instructions on lossless
		// content and still behind on sparse content.) A lane whose string is longer than 64 bits, a counter overflow
		// in reach or the partial chunk at the end of the stream take the general sink.
		const KgMasks m = kg_masks(c, base, run_start);
		if (__ballot_sync(AKOD_FULL_MASK, m.slow || (m.V | m.R) != 0) == 0)
			continue;
		uint64_t acc = 0;
		uint32_t bits = 0;
		{
			const uint4 q = *reinterpret_cast<const uint4*>(c.v);
			const uint32_t w[4] = {q.x, q.y, q.z, q.w};
			const uint32_t V = m.slow ? 0u : m.V, R = m.slow ? 0u : m.R;
			const uint32_t any_r = __reduce_or_sync(AKOD_FULL_MASK, R); // positions where some lane ends a run
#pragma unroll
			for (int j = 0; j < KG_ITEMS; j++)
			{
				const int a = (j & 1) ? ((int)w[j >> 1] >> 16) : (int)(short)w[j >> 1];
				const uint32_t vb = (V >> j) & 1u;
				const uint32_t pv = (((uint32_t)((a << 1) ^ (a >> 15)) + 1u) & 0xFFFFu) & (0u - vb);
				const uint32_t lv = kg_payload_len(pv) & (0u - vb);
				acc = (acc << lv) | pv;
				bits += lv;
				if (any_r & (1u << j))
				{
					const uint32_t rb = (R >> j) & 1u;
					const uint32_t before = c.start_mask & ((2u << j) - 1u);
					const uint32_t cj = before ? (uint32_t)j - (31u - (uint32_t)__clz(before)) : m.c0 + (uint32_t)j;
					const uint32_t pr = (cj - 1u) & (0u - rb);
					const uint32_t lr = kg_payload_len(pr) & (0u - rb);
					acc = (acc << lr) | pr;
					bits += lr;
				}
			}
		}
		if (m.slow)
			kg_emitters(c, m, n, base, run_start, [&](uint32_t p) { bits += kg_payload_len(p); });
		const uint32_t incl = warp_incl_sum(bits);
		const uint32_t step_total = __shfl_sync(AKOD_FULL_MASK, incl, 31);
		overflow = overflow || (bitpos + step_total > KG_SLOT_BITS);
		if (bits && !overflow)
		{
			const uint32_t pos = bitpos + incl - bits;
			if (bits <= 64u && !m.slow)
			{
				// MSB-first: the string's first bit is bit (31 - pos % 32) of word pos / 32
				const uint64_t x = acc << (64u - bits);
				const uint32_t xh = (uint32_t)(x >> 32), xl = (uint32_t)x, sh = pos & 31u;
				uint32_t* const at = bitbuf + (pos >> 5);
				const uint32_t w0 = xh >> sh, w1 = __funnelshift_r(xl, xh, sh), w2 = sh ? (xl << (32u - sh)) : 0u;
				atomicOr(at, w0);
				if (w1)
					atomicOr(at + 1, w1);
				if (w2)
					atomicOr(at + 2, w2);
			}
			else
			{
				KgSink sink;
				sink.open(bitbuf, pos);
				if (m.slow)
					kg_emitters(c, m, n, base, run_start, [&](uint32_t p) { sink.put(p, kg_payload_len(p)); });
				else
					kg_lane_codes(c, m, [&](uint32_t p, uint32_t len) { sink.put(p, len); });
				sink.close();
			}
		}
		bitpos += step_total;
	}
	return bitpos;
}

__device__ __noinline__ uint32_t kgl_block_tail(const int16_t* __restrict__ in, uint64_t n, uint64_t first, uint32_t carry_start,
                                                uint32_t* bitbuf, int lane)
{
	KgFetch f;
	f.q = make_uint4(0, 0, 0, 0);
	f.before = f.after = 0;
	return kgl_block_body<false>(in, n, first, carry_start, bitbuf, f, lane);
}

#ifndef KGL_CTAS
#define KGL_CTAS 5
#endif
__global__ void __launch_bounds__(KG_THREADS, KGL_CTAS)
    k_kg_lengths(const int16_t* __restrict__ in, uint64_t in_stride, uint64_t n, const uint32_t* __restrict__ blk_carry,
                 uint32_t* __restrict__ blk_bits, uint32_t nblocks, uint32_t* __restrict__ slots,
                 const uint32_t* __restrict__ blk_own, const uint8_t* __restrict__ blk_first, uint32_t blocks_per_warp)
{
	__shared__ uint32_t bitbuf_all[KGL_WARPS][KG_SLOT_WORDS + 2];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t* const bitbuf = bitbuf_all[wid];
	const uint64_t row = (uint64_t)nblocks * blockIdx.y;
	in += in_stride * blockIdx.y;
	for (uint32_t i = lane; i < KG_SLOT_WORDS + 2; i += 32)
		bitbuf[i] = 0;
	__syncwarp();

	const uint32_t b_first = (blockIdx.x * KGL_WARPS + wid) * blocks_per_warp;
	for (uint32_t b = b_first; b < b_first + blocks_per_warp && b < nblocks; b++)
	{
		const uint64_t bi = row + b;
		const bool has_next = b + 1 < nblocks;
		const uint64_t first = (uint64_t)b * KG_BLOCK;
		const bool whole = first + KG_BLOCK <= n;
		// the three per-block words of pass 1, fetched together (independent loads)
		const uint32_t own = __ldg(blk_own + bi);
		const uint32_t carry_start = __ldg(blk_carry + bi);
		const uint32_t next_first = __ldg(blk_first + (has_next ? bi + 1 : bi));
		// the first step's values are asked for before the summary words are looked at
		KgFetch f;
		if (whole)
			f = kg_fetch(in, n, first + (uint64_t)lane * KG_ITEMS, lane);
		// A block that lies entirely inside one run which also goes on behind it emits nothing, unless the run
		// counter passes 1, 2 or 65534 inside it. Pass 1 left everything needed to see that without touching the
		// stream: quantised planes are mostly such blocks.
		if (own == 0 && has_next && next_first == 0 && whole)
		{
			const uint32_t k0 = (uint32_t)first + 1u - carry_start; // position of the block's first value in its run
			uint32_t c0 = k0;
			if (k0 > 65534u)
				c0 = ((k0 - 1) % 65534u) + 1u;
			if (c0 >= 3 && c0 + (KG_BLOCK - 1) < 65534u)
			{
				if (lane == 0)
					blk_bits[bi] = 0;
				continue;
			}
		}
		const uint32_t bitpos = whole ? kgl_block_body<true>(in, n, first, carry_start, bitbuf, f, lane)
		                              : kgl_block_tail(in, n, first, carry_start, bitbuf, lane);
		if (lane == 0)
			blk_bits[bi] = bitpos;
		__syncwarp();
		// the words this block touched leave for its slot (when the string fits) and are cleared for the next block
		const uint32_t used = min((bitpos + 31) >> 5, (uint32_t)KG_SLOT_WORDS) + 1;
		if (bitpos != 0 && bitpos <= KG_SLOT_BITS)
		{
			uint32_t* slot = slots + bi * KG_SLOT_WORDS;
			const uint32_t nwords = (bitpos + 31) >> 5;
			for (uint32_t i = lane; i < nwords; i += 32)
				slot[i] = bitbuf[i];
		}
		for (uint32_t i = lane; i < used && i < KG_SLOT_WORDS + 2; i += 32)
			bitbuf[i] = 0;
		__syncwarp();
	}
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

// pass 3: pack. A CTA takes KG_PACK_PER_CTA consecutive blocks. Blocks whose bit string pass 2 left in their slot
// (nearly all) are one warp's work each: shift the slot to the block's bit offset and store, no barrier. A block
// above 8 bits per value is re-encoded by the whole CTA. (One CTA per block made this pass CTA-launch bound: most
// blocks of a quantised stream emit nothing or a few words.)
constexpr int KG_PACK_PER_CTA = KG_THREADS / 32;

__global__ void __launch_bounds__(KG_THREADS)
    k_kg_pack(const int16_t* __restrict__ in, uint64_t in_stride, uint64_t n, const uint32_t* __restrict__ blk_carry,
              const uint64_t* __restrict__ blk_off, const uint32_t* __restrict__ blk_bits, uint32_t nblocks,
              uint8_t* __restrict__ out, uint64_t out_stride, uint64_t cap_bits, const uint32_t* __restrict__ slots)
{
	__shared__ uint32_t sm_max[33];
	__shared__ uint32_t sm_sum[33];
	__shared__ uint32_t bitbuf[KG_BLOCK + 2]; // 32 bits per value at most, +1 word of misalignment, +1 spill
	__shared__ uint32_t long_total[KG_PACK_PER_CTA];
	__shared__ uint64_t long_off[KG_PACK_PER_CTA];

	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint64_t row = (uint64_t)nblocks * blockIdx.y;
	uint32_t* const image_words = reinterpret_cast<uint32_t*>(out + out_stride * blockIdx.y);

	// ---- slot path: warp w takes block b0 + w
	{
		const uint32_t b = blockIdx.x * KG_PACK_PER_CTA + wid;
		uint32_t total = 0;
		uint64_t g0 = 0;
		if (b < nblocks)
		{
			total = __ldg(blk_bits + row + b);
			g0 = __ldg(blk_off + row + b); // both loads in one round trip
		}
		const bool fits = total != 0 && g0 + total <= cap_bits; // else: nothing emitted, or the caller reports the overflow
		if (lane == 0)
		{
			long_total[wid] = (fits && total > KG_SLOT_BITS) ? total : 0u;
			long_off[wid] = g0;
		}
		if (fits && total <= KG_SLOT_BITS)
		{
			const uint32_t* slot = slots + (row + b) * KG_SLOT_WORDS;
			const uint32_t r = (uint32_t)(g0 & 31), nin = (total + 31) >> 5;
			const uint32_t nout = (r + total + 31) >> 5;
			uint32_t* dstw = image_words + (g0 >> 5);
			for (uint32_t k = lane; k < nout; k += 32)
			{
				const uint32_t lo = (k < nin) ? __ldg(slot + k) : 0u, hi = (k > 0) ? __ldg(slot + k - 1) : 0u;
				const uint32_t be = __byte_perm(__funnelshift_r(lo, hi, r), 0, 0x0123); // bytes of the file are MSB-first
				if (k == 0 || k == nout - 1)
					atomicOr(&dstw[k], be);
				else
					dstw[k] = be;
			}
		}
	}
	__syncthreads();

	// ---- long blocks (more than 8 bits per value): the whole CTA re-encodes them one after the other
	in += in_stride * blockIdx.y;
	for (int w = 0; w < KG_PACK_PER_CTA; w++)
	{
		const uint32_t total = long_total[w];
		if (total == 0)
			continue; // uniform
		const uint32_t b = blockIdx.x * KG_PACK_PER_CTA + w;
		const uint64_t g0 = long_off[w];
		const uint64_t base = (uint64_t)b * KG_BLOCK + (uint64_t)threadIdx.x * KG_ITEMS;
		const KgChunk c = kg_load(in, n, base);
		KgCode codes[KG_ITEMS];
		const uint32_t bits = kg_thread_codes<true>(c, n, base, blk_carry[row + b], sm_max, codes);

		for (int i = threadIdx.x; i < KG_BLOCK + 2; i += KG_THREADS)
			bitbuf[i] = 0;
		uint32_t check;
		const uint32_t excl = block_excl_sum(bits, sm_sum, &check); // has __syncthreads inside: bitbuf is clear after it

		if (bits)
		{
			uint32_t pos = (uint32_t)(g0 & 31) + excl; // bit position inside bitbuf
#pragma unroll
			for (int jj = 0; jj < KG_ITEMS; jj++)
			{
				const uint32_t len = codes[jj].len;
				if (len)
				{
					const uint32_t ww = pos >> 5, sh = pos & 31;
					// MSB-first: bit 'pos' of the stream is bit (31 - pos%32) of word pos/32
					const uint64_t wide = (uint64_t)codes[jj].code << (64 - sh - len);
					atomicOr(&bitbuf[ww], (uint32_t)(wide >> 32));
					if (sh + len > 32)
						atomicOr(&bitbuf[ww + 1], (uint32_t)wide);
					pos += len;
				}
			}
		}
		__syncthreads();

		uint32_t* words = image_words + (g0 >> 5);
		const uint32_t nwords = (uint32_t)(((g0 & 31) + total + 31) >> 5);
		for (uint32_t i = threadIdx.x; i < nwords; i += KG_THREADS)
		{
			const uint32_t be = __byte_perm(bitbuf[i], 0, 0x0123); // bytes of the file are MSB-first
			if (i == 0 || i == nwords - 1)
				atomicOr(&words[i], be);
			else
				words[i] = be;
		}
		__syncthreads(); // bitbuf, sm_max, sm_sum are reused by the next long block
	}
}

// ------------------------------------------------------------------------------------------------
// Single-pass encoder. The three passes above read the stream three times and compute every code twice;
// this kernel reads it once. Blocks take a ticket (so that every lower-numbered block of the image is
// already running) and two chained scans with decoupled look-back run over the blocks of an image:
//   * run starts (a max-scan): a block that contains a run start publishes it at once; a block that lies
//     entirely inside a run publishes "pass", looks back for the nearest published start and re-publishes;
//   * bit offsets (a 64-bit sum): every block publishes its own bit count, looks back, publishes the
//     inclusive count. Together with the count travel the LAST 31 BITS of the bit string, so that a block
//     can complete the 32-bit word its predecessors left unfinished: a word is written by the block that
//     holds its last bit, with plain stores -- no atomics, no pre-zeroed output.
// Flags live in the same 8-byte words as the payload (the 16-byte bit state repeats the flag in both
// halves and readers retry on a mismatch), so no fences are needed.

constexpr uint64_t KGF_NONE = 0, KGF_PART = 1, KGF_INCL = 2;
constexpr uint64_t KGF_MASK = ((uint64_t)1 << 62) - 1;

struct __align__(16) KgBitState
{
	unsigned long long a; // flag << 62 | bit count
	unsigned long long b; // flag << 62 | last (up to 31) bits of the string, right aligned
};

__device__ __forceinline__ uint64_t kgf_ld(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void kgf_st(unsigned long long* p, uint64_t v)
{
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"((unsigned long long)v) : "memory");
}

__device__ __forceinline__ void kgf_st_bits(KgBitState* p, uint64_t flag, uint64_t count, uint32_t last31)
{
	asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((unsigned long long)((flag << 62) | count)),
	             "l"((unsigned long long)((flag << 62) | last31))
	             : "memory");
}

// nearest published run start before block b; called by all 32 lanes of warp 0
__device__ __forceinline__ uint32_t kgf_lookback_run(const unsigned long long* state, long long b, int lane)
{
	long long p0 = b - 1;
	for (;;)
	{
		const long long p = p0 - lane;
		uint64_t v = KGF_INCL << 62; // before the stream: no start
		if (p >= 0)
		{
			do
				v = kgf_ld(state + p);
			while ((v >> 62) == KGF_NONE);
		}
		const uint32_t incl = __ballot_sync(AKOD_FULL_MASK, (v >> 62) == KGF_INCL);
		if (incl)
			return __shfl_sync(AKOD_FULL_MASK, (uint32_t)v, __ffs(incl) - 1);
		p0 -= 32;
	}
}

// bits before block b and the last (up to 31) of them; called by all 32 lanes of warp 0
__device__ __forceinline__ void kgf_lookback_bits(const KgBitState* state, long long b, int lane, uint64_t& prefix,
                                                  uint32_t& tail31)
{
	uint64_t sum = 0;
	uint32_t coll = 0, ncoll = 0;
	long long p0 = b - 1;
	for (;;)
	{
		const long long p = p0 - lane;
		uint64_t a = KGF_INCL << 62, bb = KGF_INCL << 62;
		if (p >= 0)
		{
			do
			{
				a = kgf_ld(&state[p].a);
				bb = kgf_ld(&state[p].b);
			} while ((a >> 62) == KGF_NONE || (a >> 62) != (bb >> 62));
		}
		const uint32_t incl = __ballot_sync(AKOD_FULL_MASK, (a >> 62) == KGF_INCL);
		const int first = incl ? __ffs(incl) - 1 : 32; // lanes 0..first are the blocks between here and the inclusive one
		uint64_t contrib = (lane <= first) ? (a & KGF_MASK) : 0;
#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			contrib += __shfl_xor_sync(AKOD_FULL_MASK, contrib, d);
		sum += contrib;
		const uint32_t nb = (uint32_t)min((unsigned long long)(a & KGF_MASK), 31ull);
		const uint32_t last = (uint32_t)bb & 0x7FFFFFFFu;
		const int upto = min(first, 31);
		for (int l = 0; l <= upto && ncoll < 31; l++)
		{
			const uint32_t nb_l = __shfl_sync(AKOD_FULL_MASK, nb, l), last_l = __shfl_sync(AKOD_FULL_MASK, last, l);
			const uint32_t take = min(nb_l, 31u - ncoll);
			coll |= (last_l & ((1u << take) - 1u)) << ncoll; // older bits go above the ones collected so far
			ncoll += take;
		}
		if (first < 32)
			break;
		p0 -= 32;
	}
	prefix = sum;
	tail31 = coll;
}

__global__ void __launch_bounds__(KG_THREADS)
    k_kg_fused(const int16_t* __restrict__ in, uint64_t in_stride, uint64_t n, unsigned long long* __restrict__ run_state,
               KgBitState* __restrict__ bit_state, uint32_t* __restrict__ ticket, uint32_t nblocks, uint8_t* __restrict__ out,
               uint64_t out_stride, uint64_t cap_bits, uint64_t* __restrict__ total_bits)
{
	__shared__ uint32_t sm_max[33];
	__shared__ uint32_t sm_sum[33];
	__shared__ uint32_t bitbuf[KG_BLOCK + 2]; // 32 bits per value at most, +2 words of slack for the peeks
	__shared__ uint32_t s_bid, s_carry, s_tail;
	__shared__ uint64_t s_prefix;

	const uint32_t img = blockIdx.y;
	const int tid = threadIdx.x, lane = tid & 31;
	if (tid == 0)
		s_bid = atomicAdd(&ticket[img], 1u);
	for (int i = tid; i < KG_BLOCK + 2; i += KG_THREADS)
		bitbuf[i] = 0;
	__syncthreads();
	const uint32_t b = s_bid;
	in += in_stride * img;
	run_state += (uint64_t)nblocks * img;
	bit_state += (uint64_t)nblocks * img;

	const uint64_t base = (uint64_t)b * KG_BLOCK + (uint64_t)tid * KG_ITEMS;
	const KgChunk c = kg_load(in, n, base);
	uint32_t blk_last;
	const uint32_t excl_start = block_excl_last_start(c.last_start, sm_max, &blk_last);

	// ---- run-start carry
	if (tid < 32)
	{
		const bool cont = __shfl_sync(AKOD_FULL_MASK, (c.start_mask & 1u) == 0, 0); // the block's first value continues a run
		if (lane == 0)
			kgf_st(&run_state[b], blk_last ? ((KGF_INCL << 62) | blk_last) : (KGF_PART << 62));
		uint32_t carry = 0;
		if (cont)
			carry = kgf_lookback_run(run_state, (long long)b, lane);
		if (lane == 0)
		{
			if (!blk_last)
				kgf_st(&run_state[b], (KGF_INCL << 62) | carry);
			s_carry = carry;
		}
	}
	__syncthreads();

	// ---- codes, packed at bit 0 of bitbuf
	KgCode codes[KG_ITEMS];
	const uint32_t bits = kg_codes<true>(c, n, base, max(s_carry, excl_start), codes);
	uint32_t total;
	const uint32_t excl = block_excl_sum(bits, sm_sum, &total);
	if (bits)
	{
		uint32_t pos = excl;
#pragma unroll
		for (int j = 0; j < KG_ITEMS; j++)
		{
			const uint32_t len = codes[j].len;
			if (len)
			{
				const uint32_t w = pos >> 5, sh = pos & 31;
				// MSB-first: bit 'pos' of the string is bit (31 - pos%32) of word pos/32
				const uint64_t wide = (uint64_t)codes[j].code << (64 - sh - len);
				atomicOr(&bitbuf[w], (uint32_t)(wide >> 32));
				if (sh + len > 32)
					atomicOr(&bitbuf[w + 1], (uint32_t)wide);
				pos += len;
			}
		}
	}
	__syncthreads();

	// ---- bit offset
	if (tid < 32)
	{
		// own last (up to 31) bits
		const uint32_t m = min(total, 31u);
		uint32_t own = 0;
		if (m)
		{
			const uint32_t pos = total - m;
			own = __funnelshift_l(bitbuf[(pos >> 5) + 1], bitbuf[pos >> 5], pos & 31) >> (32 - m);
		}
		if (lane == 0)
			kgf_st_bits(&bit_state[b], KGF_PART, total, own);
		uint64_t prefix;
		uint32_t tail31;
		kgf_lookback_bits(bit_state, (long long)b, lane, prefix, tail31);
		if (lane == 0)
		{
			const uint32_t incl_tail = (total >= 31) ? own : (((tail31 << total) | own) & 0x7FFFFFFFu);
			kgf_st_bits(&bit_state[b], KGF_INCL, prefix + total, incl_tail);
			s_prefix = prefix;
			s_tail = tail31;
			if (b == nblocks - 1)
				total_bits[img] = prefix + total;
		}
	}
	__syncthreads();

	// ---- output: the words whose last bit lies in this block (the final block also writes the unfinished one)
	const uint64_t prefix = s_prefix;
	if (prefix + total > cap_bits) // would not fit: the caller reports the failure from the bit count
		return;
	const uint32_t r = (uint32_t)(prefix & 31);
	uint32_t nwords = (r + total) >> 5;
	if (b == nblocks - 1 && ((r + total) & 31))
		nwords++;
	uint32_t* words = reinterpret_cast<uint32_t*>(out + out_stride * img) + (prefix >> 5);
	const uint32_t tailword = s_tail;
	for (uint32_t k = tid; k < nwords; k += KG_THREADS)
	{
		const uint32_t w = __funnelshift_r(bitbuf[k], k ? bitbuf[k - 1] : tailword, r);
		words[k] = __byte_perm(w, 0, 0x0123); // bytes of the file are MSB-first
	}
}
