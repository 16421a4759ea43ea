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
//     The scan runs inside ONE kernel (k_kd_decode): spans per thread while walking, a block scan, a decoupled
//     look-back between CTAs.
//
// A token expands to one value, or to (u-1) copies of the value before it. Long runs are filled cooperatively
// with 128-bit stores.
//
// Result per block: the block's size when exactly n values came out of a chain that ends in the block's last byte
// (then the reference reports the whole block consumed). Every other block -- broken input -- is decoded again by
// k_kd_sequential, the reference's own algorithm (64-bit accumulator, refill policy, codewords of up to 27 leading
// zeros, the bytes its read-ahead has fetched as the result): status and values are the reference's on any bytes.
#pragma once

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// sequential decoder, one thread per block: akoKagariDecode / akoEliasDecodeStep as they are (kagari.c:119-163,
// :301-366), 64-bit accumulator and refill policy included, so that it answers exactly what the reference answers on
// ANY bytes: codewords of 16 to 27 leading zeros are taken (value truncated to 16 bits), and the bytes it reports
// consumed are the bytes its accumulator has fetched, read-ahead included. The parallel decoder hands every block it
// does not accept to this kernel (k_kt_fill's verdict): well-formed blocks never come here, broken ones get the
// reference's status and pixels. AKO_B200_SEQ_DECODE=1 sends everything here (the device implementation the parallel
// decoder is validated against).

struct KdImage;
__device__ __forceinline__ bool kd_wants_sequential(const KdImage* info, uint32_t img);

struct KdAcc
{
	const uint8_t* cur;
	const uint8_t* end;
	uint64_t acc;
	int usage;
};

// akoEliasDecodeStep; *bits = 0 on failure
__device__ __forceinline__ uint32_t kd_reference_step(KdAcc& r, int* bits)
{
	*bits = 0;
	if (r.acc == 0 || r.usage < 32)
	{
		while (r.usage < 56 && r.cur < r.end)
		{
			r.usage += 8;
			r.acc |= (uint64_t)(*r.cur) << (64 - r.usage);
			r.cur++;
		}
		if (r.acc == 0)
			return 0;
	}
	const uint32_t top = (uint32_t)(r.acc >> 32);
	const int z = top ? __clz(top) : 32;
	const int total = 2 * z + 1;
	if (total > r.usage)
		return 0;
	*bits = total;
	const uint32_t v = (uint32_t)(r.acc >> (64 - total)) & 0xFFFFu;
	r.acc <<= total;
	r.usage -= total;
	return v;
}

__global__ void k_kd_sequential(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off,
                                const uint64_t* __restrict__ in_size, uint64_t n, int16_t* __restrict__ out_base,
                                uint64_t out_stride, uint64_t* __restrict__ result, const KdImage* info, int rescue)
{
	const uint32_t img = blockIdx.x;
	if (threadIdx.x != 0)
		return;
	if (rescue && !kd_wants_sequential(info, img))
		return;
	const uint8_t* in = in_base + in_off[img];
	const uint64_t size = in_size[img];
	int16_t* out = out_base + out_stride * img;
	KdAcc r;
	r.cur = in;
	r.end = in + size;
	r.acc = 0;
	r.usage = 0;

	uint64_t produced = 0;
	int cn = 0;
	int16_t prev = 0;
	bool ok = size > 0 && n > 0;
	while (ok && produced < n)
	{
		int bits;
		const uint32_t u = (kd_reference_step(r, &bits) - 1u) & 0xFFFFu;
		if (bits == 0)
		{
			ok = false;
			break;
		}
		const int16_t v = (int16_t)((u >> 1) ^ (0u - (u & 1))); // kagari.c:175-178
		out[produced++] = v;
		if (produced > 1 && v == prev)
		{
			if (++cn == 2)
			{
				const uint32_t len = (kd_reference_step(r, &bits) - 1u) & 0xFFFFu;
				if (bits == 0 || produced + len > n)
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
	result[img] = ok ? (uint64_t)(r.cur - in) : 0;
}

// ------------------------------------------------------------------------------------------------
// phase 1: codeword boundaries

constexpr int KD_SUB_BITS = 128;                           // bits per thread
constexpr int KD_THREADS = 256;
constexpr int KD_CTA_BITS = KD_SUB_BITS * KD_THREADS;      // 32768 bits = 4 KiB of stream per CTA
constexpr int KD_CTA_WORDS = KD_CTA_BITS / 32;
constexpr uint32_t KD_STOP = 0xFFFFFFFFu;                  // "the chain ended before this point"
constexpr uint64_t KD_STOP64 = ~(uint64_t)0;
constexpr int KD_MAX_RUNS = 4;                             // CTA-level synchronisation runs: 0 speculates, 1 starts every chunk where its
                                                           // predecessor ended, 2 proves it (3: one more round); a stream that has not
                                                           // settled by then goes to the sequential kernel

// per-image bookkeeping, device resident
struct KdImage
{
	uint64_t changed[8];   // per sync run: number of CTAs whose end moved
	uint64_t stop_pos;     // bit position right after the last codeword
	uint64_t outputs;      // values they expand to
	uint64_t sequential;   // the parallel decoder does not accept the block: the sequential kernel decides
};

__device__ __forceinline__ bool kd_needs_rescue(const KdImage* info, uint32_t img)
{
	return info[img].changed[KD_MAX_RUNS - 1] != 0;
}

__device__ __forceinline__ bool kd_wants_sequential(const KdImage* info, uint32_t img)
{
	return info[img].sequential != 0;
}

__global__ void k_kd_init(KdImage* info, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	for (int r = 0; r < 8; r++)
		info[i].changed[r] = 0;
	info[i].stop_pos = KD_STOP64;
	info[i].outputs = 0;
	info[i].sequential = 0;
}

struct KdSubState
{
	uint32_t start; // first codeword boundary of the subsequence, relative to its CTA (KD_STOP: none)
	uint32_t count; // codewords that start in it
};

// stages one CTA's worth (+ 2 words look-ahead) of the stream as big-endian words; bytes past 'size' read as zero.
// A block sits at any byte offset of the blob: words are assembled from the two aligned 32-bit words that hold them
// (two loads instead of four byte loads and their shifts); only the words at the very end of the stream go by bytes.
__device__ __forceinline__ void kd_stage_bits(uint32_t* sm, const uint8_t* __restrict__ in, uint64_t size,
                                              uint64_t first_byte)
{
	const uint32_t a = (uint32_t)((uintptr_t)(in + first_byte) & 3u);
	const uint32_t* __restrict__ wp = reinterpret_cast<const uint32_t*>(in + first_byte - a);
	for (int i = threadIdx.x; i < KD_CTA_WORDS + 2; i += KD_THREADS)
	{
		const uint64_t b = first_byte + (uint64_t)i * 4;
		uint32_t w = 0;
		if (b + 8 <= size) // both aligned words end inside the stream
		{
			const uint32_t v = __funnelshift_r(__ldg(wp + i), __ldg(wp + i + 1), 8 * a);
			w = __byte_perm(v, 0, 0x0123);
		}
		else
		{
#pragma unroll
			for (int k = 0; k < 4; k++)
				w = (w << 8) | (uint32_t)((b + k < size) ? in[b + k] : 0);
		}
		sm[i] = w;
	}
}

// walks codewords from 'start' (relative to the CTA) until the first boundary >= limit.
// Returns that boundary (or KD_STOP if the chain ends), counts codewords starting before 'limit'.
// If stop_at != nullptr and the chain ends, *stop_at receives the relative position where it ended.
// bits_left: codewords may not reach past this CTA-relative position (the end of the stream, clamped to 32 bits).
// Both words under the cursor are fetched for every codeword (two LDS off one address): carrying them in registers
// and reloading on a word crossing is a branch that some lane takes at nearly every codeword.
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
	do
	{
		const uint32_t* at = sm + (p >> 5);
		const uint32_t w = __funnelshift_l(at[1], at[0], p & 31);
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
	} while (p < limit);
	return p;
}

struct KdNoEmit
{
	__device__ __forceinline__ void operator()(uint32_t, uint32_t) const {}
};

// One synchronisation run. run == 0: every CTA assumes it starts on a boundary.
// run > 0: CTA b starts where CTA b-1 ended in the previous run (ends_prev). Only thread 0's start can differ from
// the previous run then, and Elias gamma codes resynchronise within a few codewords: thread 0 alone re-walks its
// 128 bits (six words fetched by lanes 0-5) and, when it lands on the boundary thread 1 started from, nothing else
// in the CTA changes -- the whole CTA restages and re-walks only when it does not (fast path: a 458 MB lossless
// stream spent as long in run 1 as in run 0 before).
__global__ void __launch_bounds__(KD_THREADS)
    k_kd_sync(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off,
              const uint64_t* __restrict__ in_size, uint32_t nblk, int run, const uint64_t* __restrict__ ends_prev,
              uint64_t* __restrict__ ends_new, KdSubState* __restrict__ sub, KdImage* __restrict__ info)
{
	__shared__ uint32_t sm[KD_CTA_WORDS + 2];
	__shared__ uint32_t sm_end[KD_THREADS];
	__shared__ uint32_t sm_fast;

	const uint32_t img = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
	const uint64_t size = __ldg(in_size + img);
	const uint64_t off = __ldg(in_off + img); // fetched with the size: one round trip instead of two
	const uint64_t total_bits = size * 8;
	const uint64_t cta_bit0 = (uint64_t)b * KD_CTA_BITS;
	ends_prev += (uint64_t)nblk * img;
	ends_new += (uint64_t)nblk * img;
	sub += (uint64_t)nblk * KD_THREADS * img;

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
		}
		sub[(uint64_t)b * KD_THREADS + t] = KdSubState{KD_STOP, 0};
		return;
	}
	// codewords may not cross this (relative) position
	const uint32_t bits_left = (uint32_t)min(total_bits - cta_bit0, (uint64_t)0xFFFFFFFFu);

	if (run > 0)
	{
		if (t < 32)
		{
			uint32_t new_start = 0, old_next = 0;
			uint64_t e_old = 0;
			KdSubState old0 = KdSubState{0, 0};
			if (t == 0)
			{
				const uint64_t e_prev = b ? ends_prev[b - 1] : 0;
				new_start = (e_prev == KD_STOP64) ? KD_STOP : (uint32_t)(e_prev - cta_bit0);
				old0 = sub[(uint64_t)b * KD_THREADS];
				old_next = sub[(uint64_t)b * KD_THREADS + 1].start;
				e_old = ends_prev[b];
			}
			if (t < 6) // bits [0, 192) of the CTA: the first subsequence and the look-ahead of its last codeword
			{
				const uint8_t* in = in_base + off;
				const uint64_t at = (cta_bit0 >> 3) + (uint64_t)t * 4;
				uint32_t w = 0;
#pragma unroll
				for (int k = 0; k < 4; k++)
					w = (w << 8) | (uint32_t)((at + k < size) ? in[at + k] : 0);
				sm[t] = w;
			}
			__syncwarp();
			if (t == 0)
			{
				uint32_t fast = 0;
				if (new_start == old0.start)
				{
					ends_new[b] = e_old; // same start as in the previous run: same everything
					fast = 1;
				}
				else if (new_start != KD_STOP)
				{
					uint32_t count, stop_rel;
					const uint32_t end0 = kd_walk<false>(sm, new_start, KD_SUB_BITS, bits_left, count, &stop_rel, KdNoEmit());
					if (end0 != KD_STOP && end0 == old_next)
					{
						sub[(uint64_t)b * KD_THREADS] = KdSubState{new_start, count};
						ends_new[b] = e_old;
						fast = 1;
					}
				}
				sm_fast = fast;
			}
		}
		__syncthreads();
		if (sm_fast)
			return;
		__syncthreads(); // sm[0..5] are restaged below
	}

	// where the previous CTA's chain ended in the previous run: asked for before the staging barrier
	uint64_t e_prev = 0;
	if (t == 0 && run != 0 && b != 0)
		e_prev = ends_prev[b - 1];
	kd_stage_bits(sm, in_base + off, size, cta_bit0 >> 3);
	__syncthreads();

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
	if (t == KD_THREADS - 1)
	{
		const uint64_t e = (end == KD_STOP) ? KD_STOP64 : cta_bit0 + end;
		if (run > 0 && ends_prev[b] != e)
			atomicAdd((unsigned long long*)&info[img].changed[run], 1ull);
		ends_new[b] = e;
	}
}

// ------------------------------------------------------------------------------------------------
// phase 2: classification + expansion

constexpr int KT_LONG = 16;                     // runs at least this long are filled by the whole warp

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
	// runs of up to ~128 values: element stores, every lane busy for one to four rounds (coalesced: a round is 64
	// contiguous bytes). The head / 16-byte body / tail form below costs three times the instructions on such runs and
	// only pays on long ones (measured: thresholds 8, 64, 128, 256, 512, 2048 -> 0.957, 0.910, 0.897, 0.905, 0.907,
	// 0.915 ms for the decode pass + fill of C2)
	if (count < head + 128u)
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

constexpr uint32_t KT_BIG = 2048;   // (512 and 256 measured slower: the decode pass fills such runs faster than the list round trip) runs at least this long leave the CTA: they go to a list that k_kt_fill spreads over the GPU
constexpr uint32_t KT_PIECE = 4096; // ... in pieces of at most this many values (one warp each)

// ------------------------------------------------------------------------------------------------
// The fused pass: classification and expansion straight from the bit stream, one launch after the boundary search.
//
// A CTA takes the same 4 KiB of stream as in k_kd_sync and its 256 final (start, count) pairs. There is no token
// buffer: every thread walks its 128 bits twice from shared memory,
//   walk 1  builds the span of its codewords (4-state function + the number of values per entry state) as they are
//           decoded; the first two codewords of a subsequence are compared against the previous thread's last two
//           (handed over in shared memory, from the previous CTA through its look-back record) after the walk;
//   scan    one block-wide scan of the 256 spans; the CTA's own span is published and a decoupled look-back over the
//           preceding CTAs' spans (CTA indices are tickets, so every predecessor is running or done) yields the state
//           the CTA is entered in and its first output position, with 64-bit counts;
//   walk 2  decodes again, now knowing state and position, and writes. A warp whose values fit KF_WIN collects them
//           in shared memory and stores 16-byte rows (a thread's ~38 values lie 76 bytes from its neighbour's: direct
//           2-byte stores would cost a 32-byte sector each); other warps store directly, runs of KT_LONG..KT_BIG-1
//           values are filled by the warp together, longer ones go to big_list for k_kt_fill.
// This replaces five kernels (token extraction, two token passes, the single-CTA resolve pass, the token count scan)
// and 6 bytes of HBM traffic per codeword.

constexpr uint32_t KF_WIN = 2048;              // values a warp stages
constexpr uint32_t KF_PITCH = KF_WIN + 8 + 8;  // + misalignment of its first value, rounded to whole 16-byte rows
constexpr uint64_t KF_VALID = (uint64_t)1 << 63;
constexpr uint64_t KF_COUNT_MASK = ((uint64_t)1 << 60) - 1;

// Look-back record of one CTA, one 32-byte sector. Every word validates itself (top bit), so a record needs neither a
// flag nor a fence: a reader fetches all of it together with the CTA's prefix word in ONE round trip and simply looks
// again when a word is still empty. (A CTA's span maps at most 32768 codewords of at most 65534 values: < 2^31.)
constexpr uint32_t KF_W = 0x80000000u;
struct __align__(32) KfLook
{
	uint32_t out[4];  // KF_W | values the CTA expands to, per entry state
	uint32_t map;     // KF_W | exit states
	uint32_t pad;
	uint64_t tail;    // KF_VALID | last two codewords (second to last in bits 0-15, last in bits 16-31)
};

struct KfSpan64
{
	uint32_t map;
	uint64_t out[4];
};

// a then b
__device__ __forceinline__ KfSpan64 kf_compose(const KfSpan64& a, const KfSpan64& b)
{
	KfSpan64 r;
	r.map = 0;
#pragma unroll
	for (int s = 0; s < 4; s++)
	{
		const uint32_t mid = (a.map >> (2 * s)) & 3u;
		r.map |= ((b.map >> (2 * mid)) & 3u) << (2 * s);
		r.out[s] = a.out[s] + (mid == 0 ? b.out[0] : mid == 1 ? b.out[1] : mid == 2 ? b.out[2] : b.out[3]);
	}
	return r;
}

__device__ __forceinline__ uint32_t kf_pick(const uint32_t (&o)[4], uint32_t s)
{
	return s == 0 ? o[0] : s == 1 ? o[1] : s == 2 ? o[2] : o[3];
}

// Cursor of a walk. Every codeword fetches the two words under it again (two LDS off one address): keeping them in
// registers and reloading on a word crossing was a branch that some lane of the warp takes at nearly every codeword,
// i.e. more instructions than the loads it saved.
struct KfCursor
{
	uint32_t p;
};

__device__ __forceinline__ void kf_open(const uint32_t*, uint32_t start, KfCursor& c)
{
	c.p = start;
}

// next codeword (the boundary search has validated it: at most 15 leading zeros, inside the stream)
__device__ __forceinline__ uint32_t kf_next(const uint32_t* sm, KfCursor& c)
{
	const uint32_t* at = sm + (c.p >> 5);
	const uint32_t w = __funnelshift_l(at[1], at[0], c.p & 31);
	const int top = 31 - __clz(w | 0x10000u); // position of the leading one: 31 - zeros
	const uint32_t u = w >> (2 * top - 31);
	c.p += 63 - 2 * top;
	return u;
}

// fills dst[off, off + count) of a 16-byte aligned shared array, by the 32 lanes of a warp
__device__ __forceinline__ void kf_fill_stage(int16_t* dst, uint32_t off, uint32_t count, int16_t v, int lane)
{
	int16_t* const o = dst + off;
	const uint32_t head = (8u - (off & 7u)) & 7u;
	if (count < head + 8u)
	{
		for (uint32_t i = (uint32_t)lane; i < count; i += 32)
			o[i] = v;
		return;
	}
	if ((uint32_t)lane < head)
		o[lane] = v;
	const uint32_t vv = (uint32_t)(uint16_t)v * 0x10001u;
	const uint4 q = make_uint4(vv, vv, vv, vv);
	uint4* const body = reinterpret_cast<uint4*>(o + head);
	const uint32_t nq = (count - head) >> 3, tail0 = head + (nq << 3);
	for (uint32_t i = (uint32_t)lane; i < nq; i += 32)
		body[i] = q;
	if (tail0 + (uint32_t)lane < count)
		o[tail0 + lane] = v;
}

#ifndef KF_CTAS
#define KF_CTAS 4
#endif
__global__ void __launch_bounds__(KD_THREADS, KF_CTAS)
    k_kd_decode(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off,
                const uint64_t* __restrict__ in_size, uint32_t nblk, const KdSubState* __restrict__ sub,
                const uint64_t* __restrict__ ends_final, KfLook* __restrict__ look, uint64_t* __restrict__ prefix,
                uint32_t* __restrict__ ticket, KdImage* __restrict__ info, int16_t* __restrict__ out_base,
                uint64_t out_stride, uint64_t n_values, KtRun* __restrict__ big_list, uint32_t* __restrict__ big_count,
                uint32_t big_cap)
{
	__shared__ uint32_t sm[KD_CTA_WORDS + 2];
	__shared__ KtSpan sm_scan[KD_THREADS / 32];
	__shared__ uint64_t sm_last[KD_THREADS / 32]; // KF_VALID | the last two codewords of the warp below
	__shared__ uint32_t sm_b, sm_entry;
	__shared__ uint64_t sm_base;
	__shared__ __align__(16) int16_t stage[KD_THREADS / 32][KF_PITCH];

	const uint32_t img = blockIdx.y, t = threadIdx.x;
	const int lane = t & 31, wid = t >> 5;
	if (kd_needs_rescue(info, img))
		return; // no fixed point within KD_MAX_RUNS: the sequential kernel decodes this image
	if (t == 0)
		sm_b = atomicAdd(&ticket[img], 1u);
	if (t < KD_THREADS / 32)
		sm_last[t] = 0;
	__syncthreads();
	const uint32_t b = sm_b;
	const uint64_t size = __ldg(in_size + img);
	const uint64_t off = __ldg(in_off + img);
	const uint64_t total_bits = size * 8;
	const uint64_t cta_bit0 = (uint64_t)b * KD_CTA_BITS;
	if (cta_bit0 > total_bits)
		return;
	const bool last_cta = cta_bit0 + KD_CTA_BITS > total_bits; // the next CTA has nothing
	sub += ((uint64_t)nblk * img + b) * KD_THREADS;
	look += (uint64_t)nblk * img;
	prefix += (uint64_t)nblk * img;
	int16_t* const out = out_base + out_stride * img;

	const KdSubState s = sub[t];
	// where the next subsequence starts: KD_STOP there and not here = the chain ends in this one
	uint32_t next_start;
	if (t + 1 < KD_THREADS)
		next_start = sub[t + 1].start;
	else
		next_start = (ends_final[(uint64_t)nblk * img + b] == KD_STOP64) ? KD_STOP : 0u;
	kd_stage_bits(sm, in_base + off, size, cta_bit0 >> 3);
	__syncthreads();

	// ---------------- walk 1: the span of this subsequence
	const uint32_t K = (s.start == KD_STOP) ? 0u : s.count;
	uint32_t f0 = 0, f1 = 0, p1 = 0, p2 = 0;
	uint32_t coop_mask = 0; // bit k: entered in state k, codewords 2.. hold a run of KT_LONG values or more
	KtSpan rest = kt_identity();
	{
		KfCursor c;
		kf_open(sm, (s.start != KD_STOP) ? s.start : 0u, c);
		if (K > 0)
			p1 = f0 = kf_next(sm, c);
		if (K > 1)
		{
			p2 = p1;
			p1 = f1 = kf_next(sm, c);
		}
		// codewords 2..K-1: the four entry states side by side until they agree (two codewords in a row that differ
		// from both their predecessors send every state to V), one state and one common count afterwards
		uint32_t map = rest.map, o[4] = {0, 0, 0, 0}, x = 0, common = 0;
		bool merged = false;
		for (uint32_t j = 2; j < K; j++)
		{
			const uint32_t u = kf_next(sm, c);
			const uint32_t m = kt_step_map(u == p1, u == p2);
			if (merged)
			{
				if (x == ST_R)
				{
					common += u - 1u;
					coop_mask = (u > KT_LONG) ? 0xFu : coop_mask;
				}
				else
					common += 1u;
				x = (m >> (2 * x)) & 3u;
			}
			else
			{
				uint32_t nmap = 0;
#pragma unroll
				for (int k = 0; k < 4; k++)
				{
					const uint32_t st = (map >> (2 * k)) & 3u;
					o[k] += (st == ST_R) ? (u - 1u) : 1u;
					coop_mask |= (uint32_t)(st == ST_R && u > KT_LONG) << k;
					nmap |= ((m >> (2 * st)) & 3u) << (2 * k);
				}
				map = nmap;
				merged = kt_is_const(map);
				x = map & 3u;
			}
			p2 = p1;
			p1 = u;
		}
		rest.map = merged ? x * 0x55u : map;
#pragma unroll
		for (int k = 0; k < 4; k++)
			rest.out[k] = o[k] + common;
		// the one thread whose chain ends here records where (the boundary search has left that to this kernel)
		if (s.start != KD_STOP && next_start == KD_STOP)
			info[img].stop_pos = cta_bit0 + c.p;
	}

	// ---------------- the two codewords before this subsequence: from the lane below, the warp below (shared memory,
	// point to point: no barrier) or the CTA below (its look-back record)
	const uint32_t my_last = (p2 & 0xFFFFu) | (p1 << 16);
	if (lane == 31)
	{
		if (wid == KD_THREADS / 32 - 1)
			*(volatile uint64_t*)&look[b].tail = KF_VALID | (uint64_t)my_last;
		else
			*(volatile uint64_t*)&sm_last[wid + 1] = KF_VALID | (uint64_t)my_last;
	}
	uint32_t prev_last = __shfl_up_sync(AKOD_FULL_MASK, my_last, 1);
	uint32_t q1, q2; // q1 = the codeword right before this subsequence, q2 the one before that
	if (lane == 0)
	{
		uint64_t tl = KF_VALID;
		if (wid > 0)
		{
			do
				tl = *(volatile const uint64_t*)&sm_last[wid];
			while (!(tl & KF_VALID));
		}
		else if (b != 0)
		{
			do
				tl = *(volatile const uint64_t*)&look[b - 1].tail;
			while (!(tl & KF_VALID));
		}
		prev_last = (uint32_t)tl;
	}
	q2 = prev_last & 0xFFFFu;
	q1 = prev_last >> 16;
	if (t == 0 && b == 0)
	{
		q2 = 0x10000u; // sentinels never compare equal
		q1 = 0x20000u;
	}
	KtSpan mine = rest;
	const uint32_t m0 = kt_step_map(f0 == q1, f0 == q2);
	const uint32_t m1 = kt_step_map(f1 == f0, f1 == q1);
	if (K > 0)
	{
		mine.map = 0;
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			uint32_t st = (uint32_t)k;
			uint32_t n = (st == ST_R) ? (f0 - 1u) : 1u;
			st = (m0 >> (2 * st)) & 3u;
			if (K > 1)
			{
				n += (st == ST_R) ? (f1 - 1u) : 1u;
				st = (m1 >> (2 * st)) & 3u;
			}
			n += kf_pick(rest.out, st);
			mine.map |= ((rest.map >> (2 * st)) & 3u) << (2 * k);
			mine.out[k] = n;
		}
	}

	// ---------------- block scan of the spans with ONE barrier: warps scan themselves, then fold the totals below them
	KtSpan warp_total;
	KtSpan before = kt_warp_excl_scan(mine, &warp_total);
	if (lane == 0)
		sm_scan[wid] = warp_total;
	__syncthreads();
	{
		KtSpan carry = kt_identity();
		for (int w = 0; w < wid; w++)
			carry = kt_compose(carry, sm_scan[w]);
		before = kt_compose(carry, before);
	}

	// ---------------- look-back (warp 0): entry state and first output position of this CTA
	if (wid == 0)
	{
		KtSpan total = sm_scan[0];
#pragma unroll
		for (int w = 1; w < KD_THREADS / 32; w++)
			total = kt_compose(total, sm_scan[w]);
		uint32_t entry = ST_V;
		uint64_t base = 0;
		if (b != 0)
		{
			if (lane < 5)
				*(volatile uint32_t*)(&look[b].out[0] + lane) = KF_W | (lane < 4 ? kf_pick(total.out, (uint32_t)lane) : total.map);
			KfSpan64 F; // everything between the window under examination and this CTA
			F.map = kt_identity().map;
			F.out[0] = F.out[1] = F.out[2] = F.out[3] = 0;
			for (int64_t cur = (int64_t)b - 1;; cur -= 32)
			{
				const int64_t p = cur - lane;
				KfSpan64 A;
				A.map = kt_identity().map;
				A.out[0] = A.out[1] = A.out[2] = A.out[3] = 0;
				uint64_t pv = KF_VALID; // before the first CTA: state V, nothing written
				if (p >= 0)
				{
					for (;;)
					{
						// prefix and record together: one round trip
						pv = *(volatile const uint64_t*)&prefix[p];
						uint32_t r0, r1, r2, r3;
						asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
						             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
						             : "l"(&look[p].out[0]));
						const uint32_t rm = *(volatile const uint32_t*)&look[p].map;
						if (pv & KF_VALID)
							break;
						if (r0 & r1 & r2 & r3 & rm & KF_W)
						{
							A.map = rm & 0xFFu;
							A.out[0] = r0 & ~KF_W;
							A.out[1] = r1 & ~KF_W;
							A.out[2] = r2 & ~KF_W;
							A.out[3] = r3 & ~KF_W;
							break;
						}
					}
				}
				const uint32_t pm = __ballot_sync(AKOD_FULL_MASK, (pv & KF_VALID) != 0);
				const int first = pm ? __ffs(pm) - 1 : 32;
				if (lane >= first)
				{
					A.map = kt_identity().map;
					A.out[0] = A.out[1] = A.out[2] = A.out[3] = 0;
				}
				// Fold of the window's records in CTA order (higher lanes are older and apply first). A window costs its
				// round trip plus this fold, and that must stay below the time in which 32 newer CTAs start, or the
				// look-back of every CTA grows to the depth of all resident CTAs (measured: a fifth of the kernel's warp
				// time sat at the barrier below). Records nearly always have a constant exit state: lane l is then entered
				// in lane l+1's exit state whatever came before, and the fold is one shuffle and two warp reductions.
				KfSpan64 G;
				if (first == 0)
				{
					G.map = kt_identity().map;
					G.out[0] = G.out[1] = G.out[2] = G.out[3] = 0;
				}
				else if (__all_sync(AKOD_FULL_MASK, lane >= first || kt_is_const(A.map)))
				{
					const uint32_t exit_state = A.map & 3u;
					const uint32_t entered = __shfl_down_sync(AKOD_FULL_MASK, exit_state, 1);
					const uint32_t c = (lane + 1 < first) ? (uint32_t)(entered == 0   ? A.out[0]
					                                                   : entered == 1 ? A.out[1]
					                                                   : entered == 2 ? A.out[2]
					                                                                  : A.out[3])
					                                      : 0u; // < 2^31 each
					const uint64_t sum = (uint64_t)__reduce_add_sync(AKOD_FULL_MASK, c & 0xFFFFu) +
					                     ((uint64_t)__reduce_add_sync(AKOD_FULL_MASK, c >> 16) << 16);
					G.map = __shfl_sync(AKOD_FULL_MASK, exit_state, 0) * 0x55u;
#pragma unroll
					for (int k = 0; k < 4; k++)
						G.out[k] = (uint64_t)__shfl_sync(AKOD_FULL_MASK, (uint32_t)A.out[k], first - 1) + sum; // the oldest record
				}
				else
				{
#pragma unroll
					for (int d = 1; d < 32; d <<= 1)
					{
						KfSpan64 O;
						O.map = __shfl_down_sync(AKOD_FULL_MASK, A.map, d);
#pragma unroll
						for (int k = 0; k < 4; k++)
							O.out[k] = __shfl_down_sync(AKOD_FULL_MASK, A.out[k], d);
						if (lane + d < 32)
							A = kf_compose(O, A);
					}
					G.map = __shfl_sync(AKOD_FULL_MASK, A.map, 0);
#pragma unroll
					for (int k = 0; k < 4; k++)
						G.out[k] = __shfl_sync(AKOD_FULL_MASK, A.out[k], 0);
				}
				F = kf_compose(G, F);
				if (first < 32)
				{
					const uint64_t pf = __shfl_sync(AKOD_FULL_MASK, pv, first);
					const uint32_t st = (uint32_t)(pf >> 60) & 3u;
					entry = (F.map >> (2 * st)) & 3u;
					base = (pf & KF_COUNT_MASK) + (st == 0 ? F.out[0] : st == 1 ? F.out[1] : st == 2 ? F.out[2] : F.out[3]);
					break;
				}
			}
		}
		if (lane == 0)
		{
			const uint32_t exit_state = (total.map >> (2 * entry)) & 3u;
			const uint64_t incl = (base + kf_pick(total.out, entry)) & KF_COUNT_MASK;
			*(volatile uint64_t*)&prefix[b] = KF_VALID | ((uint64_t)exit_state << 60) | incl;
			sm_entry = entry;
			sm_base = base;
			if (last_cta)
				info[img].outputs = incl;
		}
	}
	__syncthreads();

	// ---------------- walk 2: write
	const uint32_t entry = sm_entry;
	const uint64_t cta_base = sm_base;
	uint32_t state = (before.map >> (2 * entry)) & 3u;
	uint32_t rel = kf_pick(before.out, entry);           // position relative to the CTA's first value
	const uint32_t len = kf_pick(mine.out, state);       // values this thread writes
	const uint32_t w_rel0 = __shfl_sync(AKOD_FULL_MASK, rel, 0);
	const uint32_t w_len = __shfl_sync(AKOD_FULL_MASK, rel + len, 31) - w_rel0;
	const bool staged = w_len < KF_WIN; // uniform per warp; no run of KT_BIG values fits
	const uint64_t w_pos0 = cta_base + w_rel0;
	const uint32_t mis = (uint32_t)w_pos0 & 7u;
	int16_t* const wst = stage[wid];
	// does any thread of the warp meet a run that the warp should fill together? (exact: the first two codewords are
	// classified with the entry state now known, the others were recorded per entry state in walk 1)
	bool any_coop;
	{
		bool mine_coop = false;
		uint32_t st = state;
		if (K > 0)
		{
			mine_coop = (st == ST_R) && f0 > KT_LONG;
			st = (m0 >> (2 * st)) & 3u;
		}
		if (K > 1)
		{
			mine_coop |= (st == ST_R) && f1 > KT_LONG;
			st = (m1 >> (2 * st)) & 3u;
		}
		mine_coop |= ((coop_mask >> st) & 1u) != 0;
		any_coop = __any_sync(AKOD_FULL_MASK, mine_coop);
	}
	p1 = q1;
	p2 = q2;
	auto emit_literal = [&](uint32_t u) {
		if (staged)
			wst[rel - w_rel0 + mis] = kt_value(u);
		else if (cta_base + rel < n_values)
			out[cta_base + rel] = kt_value(u);
		rel += 1;
		state = kt_next(state, u == p1, u == p2);
	};
	if (!any_coop)
	{
		// every run of this warp is shorter than KT_LONG: threads go their own way
		KfCursor c;
		kf_open(sm, K ? s.start : 0u, c);
		for (uint32_t j = 0; j < K; j++)
		{
			const uint32_t u = kf_next(sm, c);
			if (state == ST_R)
			{
				// an RLE count: (u - 1) more copies of the value before it (kagari.c:342-354)
				const uint32_t count = u - 1u;
				const int16_t v = kt_value(p1);
				const uint64_t pos = cta_base + rel;
				if (staged)
				{
					int16_t* d = wst + (rel - w_rel0 + mis);
					for (uint32_t k = 0; k < count; k++)
						d[k] = v;
				}
				else if (pos + count <= n_values)
				{
					int16_t* d = out + pos;
					for (uint32_t k = 0; k < count; k++)
						d[k] = v;
				}
				rel += count;
				state = ST_A;
			}
			else
				emit_literal(u);
			p2 = p1;
			p1 = u;
		}
	}
	else
	{
		const uint32_t Kmax = __reduce_max_sync(AKOD_FULL_MASK, K);
		KfCursor c;
		kf_open(sm, K ? s.start : 0u, c);
		for (uint32_t j = 0; j < Kmax; j++)
		{
			bool coop = false;
			uint32_t count = 0;
			int16_t v = 0;
			if (j < K)
			{
				const uint32_t u = kf_next(sm, c);
				if (state == ST_R)
				{
					count = u - 1u;
					v = kt_value(p1);
					const uint64_t pos = cta_base + rel;
					const bool fits = pos + count <= n_values;
					if (staged || fits)
					{
						if (count >= KT_BIG && !staged)
						{
							// (a staged warp fills its runs in shared memory whatever their length: they fit KF_WIN)
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
							coop = true;
						else if (staged)
						{
							int16_t* d = wst + (rel - w_rel0 + mis);
							for (uint32_t k = 0; k < count; k++)
								d[k] = v;
						}
						else
						{
							int16_t* d = out + pos;
							for (uint32_t k = 0; k < count; k++)
								d[k] = v;
						}
					}
					if (!coop)
						rel += count;
					state = ST_A;
				}
				else
					emit_literal(u);
				p2 = p1;
				p1 = u;
			}
			// runs of KT_LONG .. KT_BIG-1 values: the warp fills them together
			uint32_t todo = __ballot_sync(AKOD_FULL_MASK, coop);
			while (todo)
			{
				const int src = __ffs(todo) - 1;
				todo &= todo - 1;
				const uint32_t r_rel = __shfl_sync(AKOD_FULL_MASK, rel, src);
				const uint32_t r_count = __shfl_sync(AKOD_FULL_MASK, count, src);
				const int16_t r_v = (int16_t)__shfl_sync(AKOD_FULL_MASK, (int)v, src);
				if (staged)
					kf_fill_stage(wst, r_rel - w_rel0 + mis, r_count, r_v, lane);
				else
					kt_fill_warp(out, cta_base + r_rel, r_count, r_v, lane);
			}
			if (coop)
				rel += count;
		}
	}
	if (staged)
	{
		__syncwarp();
		// stage[i] is value w_pos0 - mis + i: rows of eight are 16-byte aligned on both sides
		uint64_t room = (w_pos0 < n_values) ? n_values - w_pos0 : 0; // (a broken stream may expand past the plane)
		const uint32_t n = (uint32_t)min((uint64_t)w_len, room);
		const uint32_t lo = mis, hi = mis + n;
		const uint32_t row0 = (lo + 7) >> 3, row1 = hi >> 3;
		int16_t* const g = out + (w_pos0 - mis);
		if (row0 >= row1)
		{
			for (uint32_t i = lo + lane; i < hi; i += 32)
				g[i] = wst[i];
		}
		else
		{
			if (lo + lane < row0 * 8)
				g[lo + lane] = wst[lo + lane];
			const uint4* src = reinterpret_cast<const uint4*>(wst);
			uint4* dst = reinterpret_cast<uint4*>(g);
			for (uint32_t r = row0 + lane; r < row1; r += 32)
				dst[r] = src[r];
			if (row1 * 8 + lane < hi)
				g[row1 * 8 + lane] = wst[row1 * 8 + lane];
		}
	}
}

// the big runs, one warp per piece, spread over the whole GPU; the first thread of each image also gives the verdict.
// A well-formed block expands to exactly n values and ends in the last byte of the block: then the reference reports
// the whole block consumed. Everything else (no fixed point of the boundary search, a chain that stops early or on a
// codeword of more than 15 leading zeros, trailing bytes, too few or too many values) is left to the sequential
// kernel, which answers as the reference does.
__global__ void __launch_bounds__(256)
    k_kt_fill(const KtRun* __restrict__ big_list, const uint32_t* __restrict__ big_count, uint32_t big_cap,
              int16_t* __restrict__ out_base, uint64_t out_stride, KdImage* __restrict__ info, uint64_t n_values,
              const uint64_t* __restrict__ in_size, uint64_t* __restrict__ result)
{
	const uint32_t img = blockIdx.y;
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		const uint64_t size = in_size[img];
		const bool ok = info[img].outputs == n_values && !kd_needs_rescue(info, img) && info[img].stop_pos != KD_STOP64 &&
		                ((info[img].stop_pos + 7) >> 3) == size;
		result[img] = ok ? size : 0;
		info[img].sequential = ok ? 0 : 1;
	}
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
