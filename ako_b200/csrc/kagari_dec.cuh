// kagari_dec.cuh -- Kagari decoder. Replaces akoKagariDecode / akoEliasDecodeStep
// (reference library/kagari.c:119-163, :301-366).
#pragma once

#include "common.cuh"

// 32 bits of the MSB-first stream starting at bit 'pos' (zero beyond 'size' bytes)
__device__ __forceinline__ uint32_t kd_peek32(const uint8_t* __restrict__ in, uint64_t size, uint64_t pos)
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

// Reference-shaped sequential decoder: one thread per block. Kept as the simple, obviously-right device
// implementation the parallel decoder is validated against (AKO_B200_SEQ_DECODE=1 selects it).
__global__ void k_kd_sequential(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off,
                                const uint64_t* __restrict__ in_size, uint64_t n, int16_t* __restrict__ out_base,
                                uint64_t out_stride, uint64_t* __restrict__ result)
{
	const uint32_t img = blockIdx.x;
	if (threadIdx.x != 0)
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
		uint32_t w = kd_peek32(in, size, pos);
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
				w = kd_peek32(in, size, pos);
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
