// kagari_dec.cuh -- parallel Kagari decoder. Replaces akoKagariDecode / akoEliasDecodeStep
// (reference library/kagari.c:119-163, :301-366), which parse one monolithic bit stream sequentially.
//
// Two sequential dependencies are broken up:
//
// (1) Codeword boundaries. next(p) = p + 2*clz(bits at p) + 1. The stream is cut into 128-bit subsequences,
//     one thread each. Every thread first decodes from its subsequence start as if it were a boundary, then
//     threads repeatedly restart from their predecessor's real end until nothing changes: Elias gamma codes
//     self-synchronise after a few codewords, so this takes 2-3 rounds; it is correct for any input because
//     it only stops at the fixed point. The same is done between CTAs (32 Kibit each) by re-running the
//     kernel with the previous run's CTA ends; a run that changes nothing proves the chain consistent.
//
// (2) Which codeword is a value and which an RLE count (kagari.c:337-355). Written over the raw codeword
//     values u[j], the reference's (previous_value, consecutive_no) state is a 4-state machine whose
//     transitions only look at u[j] == u[j-1] and u[j] == u[j-2]:
//         V  : last token was a value, 0 repeats  : u[j]==u[j-1] ? S : V
//         A  : last token was an RLE count        : u[j]==u[j-2] ? S : V     (previous_value predates the count)
//         S  : 1 repeat so far                    : u[j]==u[j-1] ? R : V
//         R  : this token IS the RLE count        : -> A
//     Functions on 4 states compose associatively, so a scan classifies every token; carrying, per entry
//     state, the number of values a span of tokens expands to gives every token its output position too.
//
// A token expands to one value, or to (u-1) copies of the value before it. Long runs are filled
// cooperatively by the whole CTA with 128-bit stores.
//
// Result per block: bytes consumed = ceil(end of last codeword / 8) when exactly n values came out, else 0.
// (The reference reports what its 64-bit accumulator happened to have fetched, kagari.c:365; for every
// well-formed block both equal block_size. Blocks with trailing garbage are rejected here.)
#pragma once

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// sequential decoder, one thread per block: the simple device implementation the parallel decoder is
// validated against (AKO_B200_SEQ_DECODE=1 selects it)

__device__ __forceinline__ uint32_t kd_peek32_bytes(const uint8_t* __restrict__ in, uint64_t size, uint64_t pos)
{
	const uint64_t byte = pos >> 3;
	uint64_t acc = 0;
#pragma unroll
	for (int i = 0; i < 5; i++)
	{
		const uint64_t b = byte + i;
		acc = (acc << 8) | (uint64_t)((b < size) ? in[b] : 0);
	}
	return (uint32_t)(acc >> (8 - (pos & 7)));
}

struct KdImage;
__device__ __forceinline__ bool kd_needs_rescue(const KdImage* info, uint32_t img);

__global__ void k_kd_sequential(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off,
                                const uint64_t* __restrict__ in_size, uint64_t n, int16_t* __restrict__ out_base,
                                uint64_t out_stride, uint64_t* __restrict__ result, const KdImage* info, int rescue)
{
	const uint32_t img = blockIdx.x;
	if (threadIdx.x != 0)
		return;
	if (rescue && !kd_needs_rescue(info, img))
		return;
	const uint8_t* in = in_base + in_off[img];
	const uint64_t size = in_size[img];
	int16_t* out = out_base + out_stride * img;
	const uint64_t total_bits = size * 8;

	uint64_t pos = 0, produced = 0;
	int cn = 0;
	int16_t prev = 0;
	bool ok = size > 0 && n > 0;

	while (ok && produced < n)
	{
		uint32_t w = kd_peek32_bytes(in, size, pos);
		int z = w ? __clz(w) : 32;
		if (z > 15 || pos + 2 * z + 1 > total_bits)
		{
			ok = false;
			break;
		}
		uint32_t u = ((w >> (31 - 2 * z)) - 1) & 0xFFFFu;
		pos += 2 * z + 1;
		const int16_t v = (int16_t)((u >> 1) ^ (0u - (u & 1))); // kagari.c:175-178
		out[produced++] = v;
		if (produced > 1 && v == prev)
		{
			if (++cn == 2)
			{
				w = kd_peek32_bytes(in, size, pos);
				z = w ? __clz(w) : 32;
				if (z > 15 || pos + 2 * z + 1 > total_bits)
				{
					ok = false;
					break;
				}
				const uint32_t len = ((w >> (31 - 2 * z)) - 1) & 0xFFFFu;
				pos += 2 * z + 1;
				if (produced + len > n)
				{
					ok = false;
					break;
				}
				for (uint32_t j = 0; j < len; j++)
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
	result[img] = ok ? ((pos + 7) >> 3) : 0;
}

// ------------------------------------------------------------------------------------------------
// phase 1: codeword boundaries

constexpr int KD_SUB_BITS = 128;                           // bits per thread
constexpr int KD_THREADS = 256;
constexpr int KD_CTA_BITS = KD_SUB_BITS * KD_THREADS;      // 32768 bits = 4 KiB of stream per CTA
constexpr int KD_CTA_WORDS = KD_CTA_BITS / 32;
constexpr uint32_t KD_STOP = 0xFFFFFFFFu;                  // "the chain ended before this point"
constexpr uint64_t KD_STOP64 = ~(uint64_t)0;
constexpr int KD_MAX_RUNS = 6;                             // CTA-level synchronisation runs (later ones are no-ops once stable)

// per-image bookkeeping, device resident
struct KdImage
{
	uint64_t changed[8];   // per sync run: number of CTAs whose end moved
	uint64_t stop_pos;     // bit position right after the last codeword
	uint64_t tokens;       // number of codewords
	uint64_t outputs;      // values they expand to
	uint64_t overflow;     // token buffer too small / inconsistent
};

__device__ __forceinline__ bool kd_needs_rescue(const KdImage* info, uint32_t img)
{
	return info[img].changed[KD_MAX_RUNS - 1] != 0;
}

__global__ void k_kd_init(KdImage* info, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	for (int r = 0; r < 8; r++)
		info[i].changed[r] = 0;
	info[i].stop_pos = KD_STOP64;
	info[i].tokens = 0;
	info[i].outputs = 0;
	info[i].overflow = 0;
}

struct KdSubState
{
	uint32_t start; // first codeword boundary of the subsequence, relative to its CTA (KD_STOP: none)
	uint32_t count; // codewords that start in it
};

// stages one CTA's worth (+ 2 words look-ahead) of the stream as big-endian words; bytes past 'size' read as zero
__device__ __forceinline__ void kd_stage_bits(uint32_t* sm, const uint8_t* __restrict__ in, uint64_t size,
                                              uint64_t first_byte)
{
	for (int i = threadIdx.x; i < KD_CTA_WORDS + 2; i += KD_THREADS)
	{
		const uint64_t b = first_byte + (uint64_t)i * 4;
		uint32_t w = 0;
		if (b + 4 <= size)
			w = ((uint32_t)in[b] << 24) | ((uint32_t)in[b + 1] << 16) | ((uint32_t)in[b + 2] << 8) | (uint32_t)in[b + 3];
		else
		{
#pragma unroll
			for (int k = 0; k < 4; k++)
				w = (w << 8) | (uint32_t)((b + k < size) ? in[b + k] : 0);
		}
		sm[i] = w;
	}
}

__device__ __forceinline__ uint32_t kd_peek(const uint32_t* sm, uint32_t pos)
{
	return __funnelshift_l(sm[(pos >> 5) + 1], sm[pos >> 5], pos & 31);
}

// walks codewords from 'start' (relative to the CTA) until the first boundary >= limit.
// Returns that boundary (or KD_STOP if the chain ends), counts codewords starting before 'limit'.
// If stop_at != nullptr and the chain ends, *stop_at receives the relative position where it ended.
// bits_left: codewords may not reach past this CTA-relative position (the end of the stream, clamped to 32 bits).
// The two words under the cursor ride in registers and one shared load follows per word crossed, instead of two
// per codeword (a lossless stream averages 3.4 bits per codeword).
// EMIT: token j of the walk goes to emit(j, raw codeword value).
template <bool EMIT, typename F>
__device__ __forceinline__ uint32_t kd_walk(const uint32_t* sm, uint32_t start, uint32_t limit, uint32_t bits_left,
                                            uint32_t& count, uint32_t* stop_at, F emit)
{
	count = 0;
	if (start == KD_STOP)
		return KD_STOP;
	uint32_t p = start;
	if (p >= limit)
		return p;
	uint32_t idx = p >> 5;
	uint32_t hi = sm[idx], lo = sm[idx + 1];
	do
	{
		const uint32_t w = __funnelshift_l(lo, hi, p & 31);
		const int z = __clz(w); // 32 for w == 0
		const uint32_t len = 2 * z + 1;
		if (w < 0x10000u || p + len > bits_left) // more than 15 leading zeros, or past the end of the stream
		{
			if (stop_at)
				*stop_at = p;
			return KD_STOP;
		}
		if (EMIT)
			emit(count, w >> (31 - 2 * z));
		count++;
		p += len;
		if ((p >> 5) != idx) // len <= 31: at most one word further
		{
			idx++;
			hi = lo;
			lo = sm[idx + 1];
		}
	} while (p < limit);
	return p;
}

struct KdNoEmit
{
	__device__ __forceinline__ void operator()(uint32_t, uint32_t) const {}
};

// One synchronisation run. run == 0: every CTA assumes it starts on a boundary.
// run > 0: CTA b starts where CTA b-1 ended in the previous run (ends_prev).
__global__ void __launch_bounds__(KD_THREADS)
    k_kd_sync(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off,
              const uint64_t* __restrict__ in_size, uint32_t nblk, int run, const uint64_t* __restrict__ ends_prev,
              uint64_t* __restrict__ ends_new, KdSubState* __restrict__ sub, uint32_t* __restrict__ blk_count,
              KdImage* __restrict__ info)
{
	__shared__ uint32_t sm[KD_CTA_WORDS + 2];
	__shared__ uint32_t sm_end[KD_THREADS];
	__shared__ uint32_t sm_sum[33];

	const uint32_t img = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
	const uint64_t size = __ldg(in_size + img);
	const uint64_t off = __ldg(in_off + img); // fetched with the size: one round trip instead of two
	const uint64_t total_bits = size * 8;
	const uint64_t cta_bit0 = (uint64_t)b * KD_CTA_BITS;
	ends_prev += (uint64_t)nblk * img;
	ends_new += (uint64_t)nblk * img;
	sub += (uint64_t)nblk * KD_THREADS * img;
	blk_count += (uint64_t)nblk * img;

	if (run >= 2 && info[img].changed[run - 1] == 0)
	{
		// the previous run changed nothing: the chain is consistent, keep its results
		if (t == 0)
			ends_new[b] = ends_prev[b];
		return;
	}
	if (cta_bit0 > total_bits)
	{
		// past the end of this image's stream (the CTA starting exactly AT the end still runs: it is the one
		// that records where a chain that fills its last CTA completely stops)
		if (t == 0)
		{
			if (run > 0 && ends_prev[b] != KD_STOP64)
				atomicAdd((unsigned long long*)&info[img].changed[run], 1ull);
			ends_new[b] = KD_STOP64;
			blk_count[b] = 0;
		}
		sub[(uint64_t)b * KD_THREADS + t] = KdSubState{KD_STOP, 0};
		return;
	}
	// where the previous CTA's chain ended in the previous run: asked for before the staging barrier
	uint64_t e_prev = 0;
	if (t == 0 && run != 0 && b != 0)
		e_prev = ends_prev[b - 1];
	kd_stage_bits(sm, in_base + off, size, cta_bit0 >> 3);
	__syncthreads();

	// codewords may not cross this (relative) position
	const uint32_t bits_left = (uint32_t)min(total_bits - cta_bit0, (uint64_t)0xFFFFFFFFu);
	uint32_t start;
	if (t == 0)
	{
		if (run == 0 || b == 0)
			start = 0;
		else
			start = (e_prev == KD_STOP64) ? KD_STOP : (uint32_t)(e_prev - cta_bit0);
	}
	else
		start = t * KD_SUB_BITS;
	const uint32_t limit = (t + 1) * KD_SUB_BITS;

	uint32_t count, stop_rel = KD_STOP;
	uint32_t end = kd_walk<false>(sm, start, limit, bits_left, count, &stop_rel, KdNoEmit());
	for (;;)
	{
		sm_end[t] = end;
		__syncthreads();
		bool moved = false;
		if (t > 0)
		{
			const uint32_t real_start = sm_end[t - 1];
			if (real_start != start)
			{
				start = real_start;
				stop_rel = KD_STOP;
				end = kd_walk<false>(sm, start, limit, bits_left, count, &stop_rel, KdNoEmit());
				moved = true;
			}
		}
		if (!__syncthreads_or(moved))
			break;
	}

	sub[(uint64_t)b * KD_THREADS + t] = KdSubState{start, count};
	// the one thread of the consistent chain that ran into the end of the stream records where
	if (start != KD_STOP && end == KD_STOP)
		info[img].stop_pos = cta_bit0 + stop_rel;

	uint32_t total;
	block_excl_sum(count, sm_sum, &total);
	if (t == KD_THREADS - 1)
	{
		const uint64_t e = (end == KD_STOP) ? KD_STOP64 : cta_bit0 + end;
		if (run > 0 && ends_prev[b] != e)
			atomicAdd((unsigned long long*)&info[img].changed[run], 1ull);
		ends_new[b] = e;
	}
	if (t == 0)
		blk_count[b] = total;
}

// exclusive sum over CTAs of codeword counts (one CTA per image) -> first token index of every CTA
__global__ void __launch_bounds__(1024)
    k_kd_scan_counts(const uint32_t* __restrict__ blk_count, uint64_t* __restrict__ blk_base, uint32_t nblk,
                     KdImage* __restrict__ info)
{
	constexpr int ITEMS = 16;
	__shared__ uint32_t sm[33];
	blk_count += (uint64_t)nblk * blockIdx.x;
	blk_base += (uint64_t)nblk * blockIdx.x;
	uint64_t carry = 0;
	for (uint32_t b0 = 0; b0 < nblk; b0 += 1024 * ITEMS)
	{
		const uint32_t first = b0 + threadIdx.x * ITEMS;
		uint32_t v[ITEMS];
		uint32_t mine = 0; // <= 32768 each
#pragma unroll
		for (int i = 0; i < ITEMS; i++)
		{
			v[i] = (first + i < nblk) ? blk_count[first + i] : 0;
			mine += v[i];
		}
		uint32_t tot;
		uint64_t at = carry + block_excl_sum(mine, sm, &tot);
#pragma unroll
		for (int i = 0; i < ITEMS; i++)
		{
			if (first + i < nblk)
				blk_base[first + i] = at;
			at += v[i];
		}
		carry += tot;
	}
	if (threadIdx.x == 0)
		info[blockIdx.x].tokens = carry;
}

// re-walks every subsequence from its final start and stores the raw codeword values. A CTA's tokens are one
// contiguous piece of the token buffer: they are collected in shared memory and leave in 16-byte rows (a thread
// storing its own tokens one by one touched a 32-byte sector per 2-byte store). What does not fit the stage
// (more than KD_STAGE tokens in 32 Kibit: codewords under 3.2 bits on average) goes out directly.
constexpr uint32_t KD_STAGE = 10240;

__global__ void __launch_bounds__(KD_THREADS)
    k_kd_extract(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off,
                 const uint64_t* __restrict__ in_size, uint32_t nblk, const KdSubState* __restrict__ sub,
                 const uint64_t* __restrict__ blk_base, uint16_t* __restrict__ tokens, uint64_t token_stride,
                 uint64_t token_cap)
{
	__shared__ uint32_t sm[KD_CTA_WORDS + 2];
	__shared__ uint32_t sm_sum[33];
	__shared__ __align__(16) uint16_t stage[KD_STAGE];
	const uint32_t img = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
	const uint64_t size = __ldg(in_size + img);
	const uint64_t off = __ldg(in_off + img);
	const uint64_t tok0 = __ldg(blk_base + (uint64_t)nblk * img + b); // independent loads: one round trip
	const uint64_t total_bits = size * 8;
	const uint64_t cta_bit0 = (uint64_t)b * KD_CTA_BITS;
	if (cta_bit0 > total_bits)
		return;
	const KdSubState s = sub[((uint64_t)nblk * img + b) * KD_THREADS + t];
	kd_stage_bits(sm, in_base + off, size, cta_bit0 >> 3);
	uint32_t total;
	const uint32_t excl = block_excl_sum(s.count, sm_sum, &total); // syncs: sm is staged after it
	uint16_t* const tok = tokens + token_stride * img;
	uint32_t count;
	kd_walk<true>(sm, s.start, (t + 1) * KD_SUB_BITS, (uint32_t)min(total_bits - cta_bit0, (uint64_t)0xFFFFFFFFu), count, nullptr,
	              [&](uint32_t j, uint32_t u) {
		              const uint32_t at = excl + j;
		              if (at < KD_STAGE)
			              stage[at] = (uint16_t)u;
		              else if (tok0 + at < token_cap)
			              tok[tok0 + at] = (uint16_t)u;
	              });
	__syncthreads();
	// staged tokens [0, n) -> tok[tok0 .. tok0 + n), clipped to the buffer: 16-byte rows of the destination
	const uint64_t room = (tok0 < token_cap) ? token_cap - tok0 : 0;
	const uint32_t n = (uint32_t)min((uint64_t)min(total, KD_STAGE), room);
	const uint32_t head = min(n, (uint32_t)((8 - (tok0 & 7)) & 7)); // tokens before the first 16-byte boundary
	if (t < head)
		tok[tok0 + t] = stage[t];
	const uint32_t rows = (n - head) >> 3;
	uint4* const dst = reinterpret_cast<uint4*>(tok + tok0 + head);
	for (uint32_t r = t; r < rows; r += KD_THREADS)
	{
		const uint16_t* src = stage + head + 8 * r;
		uint4 v;
		v.x = (uint32_t)src[0] | ((uint32_t)src[1] << 16);
		v.y = (uint32_t)src[2] | ((uint32_t)src[3] << 16);
		v.z = (uint32_t)src[4] | ((uint32_t)src[5] << 16);
		v.w = (uint32_t)src[6] | ((uint32_t)src[7] << 16);
		dst[r] = v;
	}
	const uint32_t tail0 = head + 8 * rows;
	if (tail0 + t < n)
		tok[tok0 + tail0 + t] = stage[tail0 + t];
}

// ------------------------------------------------------------------------------------------------
// phase 2: classification + expansion

constexpr int KT_THREADS = 256;
constexpr int KT_ITEMS = 8;                     // tokens per thread in the expand pass: its loop body is large, 16 unrolled copies
                                                // were 80 KB of code and a fifth of its stall samples were instruction fetch
constexpr int KT_BLOCK = KT_THREADS * KT_ITEMS; // tokens per CTA (2048), the unit pass A and B describe
constexpr int KT_SPAN_ITEMS = 16;               // the span pass has a small body: more tokens per thread, fewer scan steps
constexpr int KT_SPAN_THREADS = KT_BLOCK / KT_SPAN_ITEMS;
constexpr int KT_LONG = 16;                     // runs at least this long are filled by the whole CTA

enum : uint32_t
{
	ST_V = 0, // last token was a value, no repeat pending
	ST_A = 1, // last token was an RLE count
	ST_S = 2, // one repeat seen
	ST_R = 3  // this token is an RLE count
};

// A span of tokens as a function of the state it is entered in: exit state (2 bits each, packed) and the
// number of values it expands to.
struct KtSpan
{
	uint32_t map;
	uint32_t out[4];
};

__device__ __forceinline__ KtSpan kt_identity()
{
	KtSpan s;
	s.map = (ST_V) | (ST_A << 2) | (ST_S << 4) | (ST_R << 6);
	s.out[0] = s.out[1] = s.out[2] = s.out[3] = 0;
	return s;
}

// the four transition functions as packed maps (2 bits per entry state, V A S R from the low end):
//   e1 e2    V  A  S  R
//   0  0  -> V  V  V  A    0x40
//   1  0  -> S  V  R  A    0x72
//   0  1  -> V  S  V  A    0x48
//   1  1  -> S  S  R  A    0x7A
__device__ __forceinline__ uint32_t kt_step_map(bool e1, bool e2)
{
	return e1 ? (e2 ? 0x7Au : 0x72u) : (e2 ? 0x48u : 0x40u);
}

__device__ __forceinline__ uint32_t kt_next(uint32_t state, bool e1, bool e2)
{
	return (kt_step_map(e1, e2) >> (2 * state)) & 3u;
}

// a then b
__device__ __forceinline__ KtSpan kt_compose(const KtSpan& a, const KtSpan& b)
{
	KtSpan r;
	r.map = 0;
#pragma unroll
	for (int s = 0; s < 4; s++)
	{
		const uint32_t mid = (a.map >> (2 * s)) & 3u;
		r.map |= ((b.map >> (2 * mid)) & 3u) << (2 * s);
		r.out[s] = a.out[s] + b.out[mid];
	}
	return r;
}

__device__ __forceinline__ KtSpan kt_shfl_up(const KtSpan& v, int d)
{
	KtSpan r;
	r.map = __shfl_up_sync(AKOD_FULL_MASK, v.map, d);
#pragma unroll
	for (int s = 0; s < 4; s++)
		r.out[s] = __shfl_up_sync(AKOD_FULL_MASK, v.out[s], d);
	return r;
}

// loads this thread's ITEMS tokens plus the two before them
template <int ITEMS>
__device__ __forceinline__ int kt_load(const uint16_t* __restrict__ tok, uint64_t m, uint64_t base, uint32_t u[ITEMS + 2])
{
	int valid = 0;
	u[0] = (base >= 2 && base - 2 < m) ? tok[base - 2] : 0x10000u; // sentinels never compare equal
	u[1] = (base >= 1 && base - 1 < m) ? tok[base - 1] : 0x20000u;
	if (base + ITEMS <= m)
	{
#pragma unroll
		for (int k = 0; k < ITEMS; k += 8)
		{
			const uint4 q = __ldg(reinterpret_cast<const uint4*>(tok + base + k));
			const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
			for (int i = 0; i < 4; i++)
			{
				u[k + 2 * i + 2] = w[i] & 0xFFFFu;
				u[k + 2 * i + 3] = w[i] >> 16;
			}
		}
		valid = ITEMS;
	}
	else
	{
#pragma unroll
		for (int j = 0; j < ITEMS; j++)
		{
			u[j + 2] = 0x30000u;
			if (base + j < m)
			{
				u[j + 2] = tok[base + j];
				valid = j + 1;
			}
		}
	}
	return valid;
}

// The span of a thread's ITEMS tokens. The four entry states are simulated side by side only until they agree:
// two tokens in a row that differ from both their predecessors send every state to V (see the table above), so in
// anything but a run of repeats the states have merged after three tokens and one simulation finishes the job.
template <int ITEMS>
__device__ __forceinline__ KtSpan kt_thread_span(const uint32_t u[ITEMS + 2], int valid)
{
	constexpr int PRE = (ITEMS < 3) ? ITEMS : 3;
	uint32_t st[4] = {ST_V, ST_A, ST_S, ST_R};
	uint32_t out[4] = {0, 0, 0, 0};
#pragma unroll
	for (int j = 0; j < PRE; j++)
	{
		if (j < valid)
		{
			const uint32_t m = kt_step_map(u[j + 2] == u[j + 1], u[j + 2] == u[j]);
#pragma unroll
			for (int s = 0; s < 4; s++)
			{
				out[s] += (st[s] == ST_R) ? (u[j + 2] - 1u) : 1u;
				st[s] = (m >> (2 * st[s])) & 3u;
			}
		}
	}
	if (ITEMS > PRE)
	{
		if (st[0] == st[1] && st[1] == st[2] && st[2] == st[3])
		{
			uint32_t s1 = st[0], o1 = 0;
#pragma unroll
			for (int j = PRE; j < ITEMS; j++)
			{
				if (j < valid)
				{
					o1 += (s1 == ST_R) ? (u[j + 2] - 1u) : 1u;
					s1 = kt_next(s1, u[j + 2] == u[j + 1], u[j + 2] == u[j]);
				}
			}
#pragma unroll
			for (int s = 0; s < 4; s++)
			{
				out[s] += o1;
				st[s] = s1;
			}
		}
		else
		{
#pragma unroll
			for (int j = PRE; j < ITEMS; j++)
			{
				if (j < valid)
				{
					const uint32_t m = kt_step_map(u[j + 2] == u[j + 1], u[j + 2] == u[j]);
#pragma unroll
					for (int s = 0; s < 4; s++)
					{
						out[s] += (st[s] == ST_R) ? (u[j + 2] - 1u) : 1u;
						st[s] = (m >> (2 * st[s])) & 3u;
					}
				}
			}
		}
	}
	KtSpan r;
	r.map = st[0] | (st[1] << 2) | (st[2] << 4) | (st[3] << 6);
#pragma unroll
	for (int s = 0; s < 4; s++)
		r.out[s] = out[s];
	return r;
}

// a span whose exit state does not depend on the entry state
__device__ __forceinline__ bool kt_is_const(uint32_t map)
{
	return ((map ^ (map >> 2)) & 0x3Fu) == 0;
}

// Exclusive scan of the 32 spans of a warp; *total = all 32 composed. When every lane's span has a constant exit
// state (the rule, see kt_thread_span) lane l is entered in lane l-1's exit state whatever came before, and the scan
// is one shuffle for the states and an integer prefix sum for the counts; otherwise the generic composition runs.
// Lanes >= n_active hold nothing (the caller passes identities there).
__device__ __forceinline__ KtSpan kt_warp_excl_scan(const KtSpan& v, KtSpan* total, int n_active = 32)
{
	const int lane = threadIdx.x & 31;
	if (__all_sync(AKOD_FULL_MASK, kt_is_const(v.map) || lane >= n_active))
	{
		const uint32_t exit_state = v.map & 3u;
		const uint32_t prev_exit = __shfl_up_sync(AKOD_FULL_MASK, exit_state, 1);
		// what lane l >= 1 contributes, entered in lane l-1's exit state
		const uint32_t c = (lane == 0 || lane >= n_active)
		                       ? 0u
		                       : (prev_exit == 0 ? v.out[0] : prev_exit == 1 ? v.out[1] : prev_exit == 2 ? v.out[2] : v.out[3]);
		const uint32_t incl = warp_incl_sum(c);
		const uint32_t before = incl - c; // lanes 1 .. l-1
		const uint32_t all = __shfl_sync(AKOD_FULL_MASK, incl, 31);
		const uint32_t last_exit = __shfl_sync(AKOD_FULL_MASK, exit_state, n_active - 1);
		KtSpan r;
		r.map = (lane == 0) ? kt_identity().map : prev_exit * 0x55u;
		total->map = last_exit * 0x55u;
#pragma unroll
		for (int s = 0; s < 4; s++)
		{
			const uint32_t first = __shfl_sync(AKOD_FULL_MASK, v.out[s], 0); // lane 0 entered in state s
			r.out[s] = (lane == 0) ? 0u : first + before;
			total->out[s] = first + all;
		}
		return r;
	}
	KtSpan incl = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const KtSpan o = kt_shfl_up(incl, d);
		if (lane >= d)
			incl = kt_compose(o, incl);
	}
	KtSpan excl = kt_shfl_up(incl, 1);
	if (lane == 0)
		excl = kt_identity();
	total->map = __shfl_sync(AKOD_FULL_MASK, incl.map, 31);
#pragma unroll
	for (int s = 0; s < 4; s++)
		total->out[s] = __shfl_sync(AKOD_FULL_MASK, incl.out[s], 31);
	return excl;
}

// Block-wide exclusive scan of spans. sm must hold 33 KtSpan. Returns the span of everything before this thread;
// *total = span of the whole CTA.
__device__ __forceinline__ KtSpan kt_block_excl_scan(KtSpan v, KtSpan* sm, KtSpan* total)
{
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	KtSpan warp_total;
	const KtSpan excl = kt_warp_excl_scan(v, &warp_total);
	if (lane == 0)
		sm[wid] = warp_total;
	__syncthreads();
	if (wid == 0)
	{
		const KtSpan w = (lane < nw) ? sm[lane] : kt_identity();
		KtSpan all;
		const KtSpan we = kt_warp_excl_scan(w, &all, nw);
		sm[lane] = we;
		if (lane == 0)
			sm[32] = all;
	}
	__syncthreads();
	const KtSpan r = kt_compose(sm[wid], excl);
	*total = sm[32];
	__syncthreads();
	return r;
}

// pass A: the span of every block of KT_BLOCK tokens. The grid is sized for the most tokens a stream of this length
// could hold, a quantised image has a small fraction of that: CTAs walk the blocks that exist with a grid stride
// (one CTA per possible block was ~10 000 CTAs per image that only found out they had nothing to do).
__global__ void __launch_bounds__(KT_SPAN_THREADS)
    k_kt_spans(const uint16_t* __restrict__ tokens, uint64_t token_stride, uint64_t token_cap,
               const KdImage* __restrict__ info, KtSpan* __restrict__ blk_span, uint32_t nblk)
{
	__shared__ KtSpan sm[33];
	const uint32_t img = blockIdx.y;
	const uint64_t m = min(info[img].tokens, token_cap);
	const uint32_t used = (uint32_t)((m + KT_BLOCK - 1) / KT_BLOCK);
	for (uint32_t blk = blockIdx.x; blk < used; blk += gridDim.x)
	{
		const uint64_t base = (uint64_t)blk * KT_BLOCK + (uint64_t)threadIdx.x * KT_SPAN_ITEMS;
		uint32_t u[KT_SPAN_ITEMS + 2];
		const int valid = (base < m) ? kt_load<KT_SPAN_ITEMS>(tokens + token_stride * img, m, base, u) : 0;
		KtSpan total;
		kt_block_excl_scan(kt_thread_span<KT_SPAN_ITEMS>(u, valid), sm, &total);
		if (threadIdx.x == 0)
			blk_span[(uint64_t)nblk * img + blk] = total;
	}
}

// pass B: one CTA per image resolves every expand-CTA's entry state and output base, and validates the block.
// Three levels: every thread composes the spans of its own chunk of CTAs; warp 0 chains the 1024 thread
// composites (32 per lane, then the 32 lanes in order) with 64-bit output counts; every thread then walks its
// chunk again from its resolved (state, base). A 458 MB lossless block has ~150 000 spans.
constexpr int KR_THREADS = 1024;

__global__ void __launch_bounds__(KR_THREADS)
    k_kt_resolve(const KtSpan* __restrict__ blk_span, uint32_t nblk, uint32_t* __restrict__ blk_state,
                 uint64_t* __restrict__ blk_out, KdImage* __restrict__ info, uint64_t n_values, uint64_t token_cap,
                 uint64_t* __restrict__ result)
{
	__shared__ uint32_t s_map[KR_THREADS];
	__shared__ uint32_t s_out[KR_THREADS][4];
	__shared__ uint32_t s_state[KR_THREADS];
	__shared__ uint64_t s_base[KR_THREADS];
	__shared__ uint64_t s_grand;

	const uint32_t img = blockIdx.x, t = threadIdx.x;
	blk_span += (uint64_t)nblk * img;
	blk_state += (uint64_t)nblk * img;
	blk_out += (uint64_t)nblk * img;
	const uint64_t m = min(info[img].tokens, token_cap);
	const uint32_t used = (uint32_t)((m + KT_BLOCK - 1) / KT_BLOCK);
	const uint32_t chunk = (used + KR_THREADS - 1) / KR_THREADS;
	const uint32_t b0 = min(t * chunk, used), b1 = min(b0 + chunk, used);

	// a thread's chunk expands to fewer than 2^32 values for any stream that can be valid (n_values < 2^32)
	// (the loads of eight spans are in flight together: one at a time this loop was a chain of L2 round trips)
	KtSpan mine = kt_identity();
	for (uint32_t b = b0; b < b1; b += 8)
	{
		KtSpan sp[8];
#pragma unroll
		for (int i = 0; i < 8; i++)
			sp[i] = (b + i < b1) ? blk_span[b + i] : kt_identity();
#pragma unroll
		for (int i = 0; i < 8; i++)
			mine = kt_compose(mine, sp[i]);
	}
	s_map[t] = mine.map;
#pragma unroll
	for (int k = 0; k < 4; k++)
		s_out[t][k] = mine.out[k];
	__syncthreads();

	if (t < 32)
	{
		// lane composite over its 32 thread composites, 64-bit counts
		uint32_t lmap = (ST_V) | (ST_A << 2) | (ST_S << 4) | (ST_R << 6);
		uint64_t lout[4] = {0, 0, 0, 0};
		for (int i = 0; i < 32; i++)
		{
			const uint32_t c = t * 32 + i, cmap = s_map[c];
			uint32_t nmap = 0;
#pragma unroll
			for (int st = 0; st < 4; st++)
			{
				const uint32_t mid = (lmap >> (2 * st)) & 3u;
				nmap |= ((cmap >> (2 * mid)) & 3u) << (2 * st);
				lout[st] += s_out[c][mid];
			}
			lmap = nmap;
		}
		// entry (state, base) of each lane: the 32 lanes in order
		uint32_t state = ST_V;
		uint64_t base = 0;
		for (int l = 0; l < 32; l++)
		{
			const uint32_t map_l = __shfl_sync(AKOD_FULL_MASK, lmap, l);
			const uint64_t o0 = __shfl_sync(AKOD_FULL_MASK, lout[0], l), o1 = __shfl_sync(AKOD_FULL_MASK, lout[1], l);
			const uint64_t o2 = __shfl_sync(AKOD_FULL_MASK, lout[2], l), o3 = __shfl_sync(AKOD_FULL_MASK, lout[3], l);
			if ((int)t > l)
			{
				base += state == 0 ? o0 : state == 1 ? o1 : state == 2 ? o2 : o3;
				state = (map_l >> (2 * state)) & 3u;
			}
		}
		// entry (state, base) of each of the lane's 32 threads
		for (int i = 0; i < 32; i++)
		{
			const uint32_t c = t * 32 + i;
			s_state[c] = state;
			s_base[c] = base;
			base += s_out[c][state];
			state = (s_map[c] >> (2 * state)) & 3u;
		}
		if (t == 31)
			s_grand = base;
	}
	__syncthreads();

	uint32_t state = s_state[t];
	uint64_t base = s_base[t];
	for (uint32_t b = b0; b < b1; b += 8)
	{
		KtSpan sp[8];
#pragma unroll
		for (int i = 0; i < 8; i++)
			sp[i] = (b + i < b1) ? blk_span[b + i] : kt_identity();
#pragma unroll
		for (int i = 0; i < 8; i++)
			if (b + i < b1)
			{
				blk_state[b + i] = state;
				blk_out[b + i] = base;
				base += state == 0 ? sp[i].out[0] : state == 1 ? sp[i].out[1] : state == 2 ? sp[i].out[2] : sp[i].out[3];
				state = (sp[i].map >> (2 * state)) & 3u;
			}
	}
	if (t == 0)
	{
		const uint64_t grand = s_grand;
		info[img].outputs = grand;
		const bool ok = grand == n_values && info[img].tokens <= token_cap && !kd_needs_rescue(info, img) &&
		                info[img].stop_pos != KD_STOP64;
		result[img] = ok ? ((info[img].stop_pos + 7) >> 3) : 0;
	}
}

__device__ __forceinline__ int16_t kt_value(uint32_t u)
{
	const uint32_t z = (u - 1u) & 0xFFFFu;
	return (int16_t)((z >> 1) ^ (0u - (z & 1u))); // kagari.c:175-178
}

// fills out[pos, pos+count) with v; called by the 32 lanes of one warp (runs are spread over the CTA's warps)
__device__ __forceinline__ void kt_fill_warp(int16_t* __restrict__ out, uint64_t pos, uint32_t count, int16_t v, int lane)
{
	// one 64-bit address, everything else in 32 bits relative to it (count <= 65534 from the queue, <= KT_PIECE from
	// the big list); 'out' is 16-byte aligned, so the misalignment of the run is that of pos
	int16_t* const o = out + pos;
	const uint32_t head = (8u - ((uint32_t)pos & 7u)) & 7u; // elements before the first 16-byte boundary
	if (count < head + 8u)
	{
		for (uint32_t i = (uint32_t)lane; i < count; i += 32)
			o[i] = v;
		return;
	}
	if ((uint32_t)lane < head)
		o[lane] = v; // at most 7 head elements
	const uint32_t vv = (uint32_t)(uint16_t)v * 0x10001u;
	const uint4 q = make_uint4(vv, vv, vv, vv);
	uint4* const body = reinterpret_cast<uint4*>(o + head);
	const uint32_t nq = (count - head) >> 3, tail0 = head + (nq << 3);
	for (uint32_t i = (uint32_t)lane; i < nq; i += 32)
		body[i] = q;
	if (tail0 + (uint32_t)lane < count)
		o[tail0 + lane] = v; // at most 7 tail elements
}

struct KtRun
{
	uint64_t pos;
	uint32_t count;
	int16_t value;
};

constexpr uint32_t KT_BIG = 2048;   // runs at least this long leave the CTA: they go to a list that k_kt_fill spreads over the GPU
constexpr uint32_t KT_PIECE = 4096; // ... in pieces of at most this many values (one warp each)

// pass C: classify every token and write what it expands to. Quantised planes are mostly a few very long
// runs; whichever CTA meets their tokens would have to write megabytes alone, so those runs are only
// recorded here (big_list, at most n_values / KT_BIG pieces per image) and written by k_kt_fill.
// (Collecting a block's values in shared memory and storing them in 16-byte rows was tried: the kernel is bound by
// its instruction count, not by its 2-byte stores, and the extra barriers and registers cost 10-20 %.)
__global__ void __launch_bounds__(KT_THREADS, 5)
    k_kt_expand(const uint16_t* __restrict__ tokens, uint64_t token_stride, uint64_t token_cap,
                const KdImage* __restrict__ info, const uint32_t* __restrict__ blk_state,
                const uint64_t* __restrict__ blk_out, uint32_t nblk, int16_t* __restrict__ out_base, uint64_t out_stride,
                uint64_t n_values, KtRun* __restrict__ big_list, uint32_t* __restrict__ big_count, uint32_t big_cap)
{
	__shared__ KtSpan sm[33];
	__shared__ KtRun queue[KT_BLOCK / 3 + 1]; // an RLE count follows at least two value tokens
	__shared__ uint32_t queue_len;

	const uint32_t img = blockIdx.y;
	const uint64_t m_all = __ldg(&info[img].tokens);
	const uint64_t m = min(m_all, token_cap);
	const uint32_t used = (uint32_t)((m + KT_BLOCK - 1) / KT_BLOCK);
	int16_t* out = out_base + out_stride * img;
	// grid stride over the token blocks that exist (see k_kt_spans)
	for (uint32_t blk = blockIdx.x; blk < used; blk += gridDim.x)
	{
		// the per-block words are fetched together: behind the scan's barriers they would be a second round trip
		const uint32_t entry = __ldg(blk_state + (uint64_t)nblk * img + blk);
		const uint64_t out0 = __ldg(blk_out + (uint64_t)nblk * img + blk);
		const uint64_t cta_base = (uint64_t)blk * KT_BLOCK;
		const uint64_t base = cta_base + (uint64_t)threadIdx.x * KT_ITEMS;
		uint32_t u[KT_ITEMS + 2];
		const int valid = (base < m) ? kt_load<KT_ITEMS>(tokens + token_stride * img, m, base, u) : 0;
		if (threadIdx.x == 0)
			queue_len = 0;

		KtSpan total;
		const KtSpan before = kt_block_excl_scan(kt_thread_span<KT_ITEMS>(u, valid), sm, &total);
		uint32_t state = (before.map >> (2 * entry)) & 3u;
		uint64_t pos = out0 + before.out[entry];

#pragma unroll
		for (int j = 0; j < KT_ITEMS; j++)
		{
			if (j < valid)
			{
				const uint32_t cur = u[j + 2];
				if (state == ST_R)
				{
					// an RLE count: (cur - 1) more copies of the value before it (kagari.c:342-354)
					const uint32_t count = cur - 1u;
					const int16_t v = kt_value(u[j + 1]);
					if (pos + count <= n_values)
					{
						if (count >= KT_BIG)
						{
							const uint32_t pieces = (count + KT_PIECE - 1) / KT_PIECE;
							const uint32_t first = atomicAdd(&big_count[img], pieces);
							for (uint32_t k = 0; k < pieces; k++)
								if (first + k < big_cap) // cannot overflow for a stream that expands to n_values
								{
									KtRun r;
									r.pos = pos + (uint64_t)k * KT_PIECE;
									r.count = min(KT_PIECE, count - k * KT_PIECE);
									r.value = v;
									big_list[(uint64_t)big_cap * img + first + k] = r;
								}
						}
						else if (count >= KT_LONG)
						{
							const uint32_t slot = atomicAdd(&queue_len, 1u);
							queue[slot].pos = pos;
							queue[slot].count = count;
							queue[slot].value = v;
						}
						else
							for (uint32_t k = 0; k < count; k++)
								out[pos + k] = v;
					}
					pos += count;
					state = ST_A;
				}
				else
				{
					if (pos < n_values)
						out[pos] = kt_value(cur);
					pos += 1;
					state = kt_next(state, cur == u[j + 1], cur == u[j]);
				}
			}
		}
		__syncthreads();
		const uint32_t nq = queue_len;
		const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
		for (uint32_t i = wid; i < nq; i += KT_THREADS / 32)
			kt_fill_warp(out, queue[i].pos, queue[i].count, queue[i].value, lane);
		__syncthreads(); // the queue is reused by the next block
	}
}

// pass D: the big runs, one warp per piece, spread over the whole GPU
__global__ void __launch_bounds__(256)
    k_kt_fill(const KtRun* __restrict__ big_list, const uint32_t* __restrict__ big_count, uint32_t big_cap,
              int16_t* __restrict__ out_base, uint64_t out_stride)
{
	const uint32_t img = blockIdx.y;
	const uint32_t n = min(big_count[img], big_cap);
	const uint32_t warps = gridDim.x * (blockDim.x >> 5);
	const int lane = threadIdx.x & 31;
	int16_t* out = out_base + out_stride * img;
	for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps)
	{
		const KtRun r = big_list[(uint64_t)big_cap * img + i];
		kt_fill_warp(out, r.pos, r.count, r.value, lane);
	}
}
