/* png_min.c -- see png_min.h. PNG (ISO/IEC 15948) subset: colour types 0, 2, 4, 6 at 8 bits, non-interlaced. */
#include "png_min.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static const uint8_t SIGNATURE[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

static uint32_t be32(const uint8_t* p)
{
	return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

static void put_be32(uint8_t* p, uint32_t v)
{
	p[0] = (uint8_t)(v >> 24);
	p[1] = (uint8_t)(v >> 16);
	p[2] = (uint8_t)(v >> 8);
	p[3] = (uint8_t)v;
}

static int fail(char* err, size_t err_len, const char* what)
{
	if (err != NULL && err_len > 0)
		snprintf(err, err_len, "%s", what);
	return 1;
}

static int paeth(int a, int b, int c)
{
	const int p = a + b - c;
	const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
	return (pa <= pb && pa <= pc) ? a : (pb <= pc) ? b : c;
}

static int channels_of(int color_type)
{
	switch (color_type)
	{
	case 0: return 1;
	case 4: return 2;
	case 2: return 3;
	case 6: return 4;
	default: return 0;
	}
}

int png_min_read(const char* path, uint8_t** out_pixels, size_t* out_w, size_t* out_h, size_t* out_channels, char* err,
                 size_t err_len)
{
	uint8_t *file = NULL, *idat = NULL, *raw = NULL, *pixels = NULL;
	size_t file_size = 0, idat_size = 0;
	int rc = 1;

	{
		FILE* fp = fopen(path, "rb");
		if (fp == NULL)
			return fail(err, err_len, "failed to open file for reading");
		fseek(fp, 0, SEEK_END);
		const long end = ftell(fp);
		fseek(fp, 0, SEEK_SET);
		if (end <= 0 || (file = malloc((size_t)end)) == NULL || fread(file, 1, (size_t)end, fp) != (size_t)end)
		{
			fclose(fp);
			free(file);
			return fail(err, err_len, "failed to read file");
		}
		fclose(fp);
		file_size = (size_t)end;
	}

	if (file_size < 8 + 25 || memcmp(file, SIGNATURE, 8) != 0)
	{
		fail(err, err_len, "incorrect PNG signature, it's no PNG or corrupted");
		goto done;
	}

	size_t w = 0, h = 0, channels = 0;
	int seen_head = 0, seen_end = 0;
	if ((idat = malloc(file_size)) == NULL)
	{
		fail(err, err_len, "out of memory");
		goto done;
	}
	for (size_t pos = 8; pos + 12 <= file_size && !seen_end;)
	{
		const uint32_t len = be32(file + pos);
		const uint8_t* type = file + pos + 4;
		const uint8_t* data = file + pos + 8;
		if ((size_t)len > file_size - pos - 12)
		{
			fail(err, err_len, "chunk length larger than the file");
			goto done;
		}
		if (be32(data + len) != (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4))
		{
			fail(err, err_len, "chunk CRC mismatch");
			goto done;
		}
		if (memcmp(type, "IHDR", 4) == 0)
		{
			if (len != 13)
			{
				fail(err, err_len, "invalid IHDR chunk");
				goto done;
			}
			w = be32(data);
			h = be32(data + 4);
			const int depth = data[8], color_type = data[9], interlace = data[12];
			channels = (size_t)channels_of(color_type);
			if (w == 0 || h == 0)
			{
				fail(err, err_len, "image with zero dimensions");
				goto done;
			}
			if (channels == 0)
			{
				char msg[64];
				snprintf(msg, sizeof(msg), "Unsupported channels number (%i)", color_type); /* akoenc.cpp:86 */
				fail(err, err_len, msg);
				goto done;
			}
			if (depth != 8)
			{
				char msg[64];
				snprintf(msg, sizeof(msg), "Unsupported bits per pixel-component (%i)", depth); /* akoenc.cpp:90 */
				fail(err, err_len, msg);
				goto done;
			}
			if (interlace != 0)
			{
				fail(err, err_len, "interlaced PNG files are not supported by this tool");
				goto done;
			}
			seen_head = 1;
		}
		else if (memcmp(type, "IDAT", 4) == 0)
		{
			memcpy(idat + idat_size, data, len);
			idat_size += len;
		}
		else if (memcmp(type, "IEND", 4) == 0)
			seen_end = 1;
		pos += 12 + (size_t)len;
	}
	if (!seen_head || idat_size == 0)
	{
		fail(err, err_len, "no IHDR / IDAT chunk");
		goto done;
	}

	/* sizes from an untrusted IHDR: refuse what would wrap (w = h = 2^31 ...) before anything is allocated */
	size_t row, raw_bytes, image_bytes;
	if (w == 0 || h == 0 || __builtin_mul_overflow(w, channels, &row) || row > ((size_t)1 << 40) ||
	    __builtin_mul_overflow(row + 1, h, &raw_bytes) || __builtin_mul_overflow(row, h, &image_bytes) ||
	    raw_bytes > ((size_t)1 << 40))
	{
		fail(err, err_len, "image dimensions out of range");
		goto done;
	}
	uLongf raw_size = (uLongf)raw_bytes;
	if ((raw = malloc(raw_size)) == NULL || (pixels = malloc(image_bytes)) == NULL)
	{
		fail(err, err_len, "out of memory");
		goto done;
	}
	if (uncompress(raw, &raw_size, idat, (uLong)idat_size) != Z_OK || raw_size != (uLongf)raw_bytes)
	{
		fail(err, err_len, "corrupted image data");
		goto done;
	}

	for (size_t y = 0; y < h; y++)
	{
		const uint8_t* src = raw + y * (row + 1);
		uint8_t* dst = pixels + y * row;
		const uint8_t* up = (y > 0) ? dst - row : NULL;
		const int filter = src[0];
		src++;
		for (size_t x = 0; x < row; x++)
		{
			const int a = (x >= channels) ? dst[x - channels] : 0;
			const int b = (up != NULL) ? up[x] : 0;
			const int c = (up != NULL && x >= channels) ? up[x - channels] : 0;
			int pred;
			switch (filter)
			{
			case 0: pred = 0; break;
			case 1: pred = a; break;
			case 2: pred = b; break;
			case 3: pred = (a + b) / 2; break;
			case 4: pred = paeth(a, b, c); break;
			default: fail(err, err_len, "unknown filter type"); goto done;
			}
			dst[x] = (uint8_t)(src[x] + pred);
		}
	}

	*out_pixels = pixels;
	pixels = NULL;
	*out_w = w;
	*out_h = h;
	*out_channels = channels;
	rc = 0;

done:
	free(file);
	free(idat);
	free(raw);
	free(pixels);
	return rc;
}

static size_t put_chunk(uint8_t* out, const char* type, const uint8_t* data, size_t len)
{
	put_be32(out, (uint32_t)len);
	memcpy(out + 4, type, 4);
	if (len > 0)
		memcpy(out + 8, data, len);
	put_be32(out + 8 + len, (uint32_t)crc32(crc32(0L, Z_NULL, 0), out + 4, (uInt)(len + 4)));
	return len + 12;
}

static void filter_row(int filter, const uint8_t* cur, const uint8_t* up, size_t row, size_t channels, uint8_t* out)
{
	for (size_t x = 0; x < row; x++)
	{
		const int a = (x >= channels) ? cur[x - channels] : 0;
		const int b = (up != NULL) ? up[x] : 0;
		const int c = (up != NULL && x >= channels) ? up[x - channels] : 0;
		int pred = 0;
		switch (filter)
		{
		case 1: pred = a; break;
		case 2: pred = b; break;
		case 3: pred = (a + b) / 2; break;
		case 4: pred = paeth(a, b, c); break;
		default: break;
		}
		out[x] = (uint8_t)(cur[x] - pred);
	}
}

int png_min_encode(const uint8_t* pixels, size_t w, size_t h, size_t channels, int effort, uint8_t** out,
                   size_t* out_size, char* err, size_t err_len)
{
	static const int COLOR_TYPE[5] = {0, 0, 4, 2, 6};
	if (channels < 1 || channels > 4)
	{
		char msg[64];
		snprintf(msg, sizeof(msg), "Unsupported channels number (%zu)", channels); /* akodec.cpp:199 */
		return fail(err, err_len, msg);
	}
	if (w == 0 || h == 0 || w > 0x7fffffffu || h > 0x7fffffffu)
		return fail(err, err_len, "invalid image dimensions");
	effort = (effort < 1) ? 1 : (effort > 10) ? 10 : effort;

	const size_t row = w * channels;
	uint8_t* raw = malloc((row + 1) * h);
	uint8_t* trial = malloc(row);
	if (raw == NULL || trial == NULL)
	{
		free(raw);
		free(trial);
		return fail(err, err_len, "out of memory");
	}

	/* effort 1: no filtering; otherwise the filter with the smallest sum of absolute residuals per row */
	for (size_t y = 0; y < h; y++)
	{
		const uint8_t* cur = pixels + y * row;
		const uint8_t* up = (y > 0) ? cur - row : NULL;
		uint8_t* dst = raw + y * (row + 1);
		int best = 0;
		if (effort > 1)
		{
			uint64_t best_sum = UINT64_MAX;
			for (int f = 0; f < 5; f++)
			{
				filter_row(f, cur, up, row, channels, trial);
				uint64_t sum = 0;
				for (size_t x = 0; x < row; x++)
					sum += (trial[x] < 128) ? trial[x] : 256u - trial[x];
				if (sum < best_sum)
				{
					best_sum = sum;
					best = f;
				}
			}
		}
		dst[0] = (uint8_t)best;
		filter_row(best, cur, up, row, channels, dst + 1);
	}
	free(trial);

	uLongf z_size = compressBound((uLong)((row + 1) * h));
	uint8_t* z = malloc(z_size);
	uint8_t* file = NULL;
	if (z == NULL || compress2(z, &z_size, raw, (uLong)((row + 1) * h), (effort >= 9) ? 9 : effort) != Z_OK ||
	    (file = malloc(8 + 25 + 12 + z_size + 12)) == NULL)
	{
		free(raw);
		free(z);
		return fail(err, err_len, "compression failed");
	}
	free(raw);

	size_t pos = 0;
	memcpy(file, SIGNATURE, 8);
	pos += 8;
	uint8_t head[13];
	put_be32(head, (uint32_t)w);
	put_be32(head + 4, (uint32_t)h);
	head[8] = 8;
	head[9] = (uint8_t)COLOR_TYPE[channels];
	head[10] = head[11] = head[12] = 0;
	pos += put_chunk(file + pos, "IHDR", head, 13);
	pos += put_chunk(file + pos, "IDAT", z, z_size);
	pos += put_chunk(file + pos, "IEND", NULL, 0);
	free(z);

	*out = file;
	*out_size = pos;
	return 0;
}

int png_min_write(const char* path, const uint8_t* pixels, size_t w, size_t h, size_t channels, int effort, char* err,
                  size_t err_len)
{
	uint8_t* file;
	size_t size;
	if (png_min_encode(pixels, w, h, channels, effort, &file, &size, err, err_len) != 0)
		return 1;
	FILE* fp = fopen(path, "wb");
	if (fp == NULL || fwrite(file, 1, size, fp) != size)
	{
		if (fp != NULL)
			fclose(fp);
		free(file);
		return fail(err, err_len, "Write error");
	}
	fclose(fp);
	free(file);
	return 0;
}
