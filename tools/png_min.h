/*
 * png_min.h -- the PNG subset the Ako tools need, on top of zlib.
 *
 * The reference tools use lodepng for this (tools/akoenc.cpp:57-97, tools/akodec.cpp:186-212), which is third-party
 * code and stays out of this repository. What the tools accept is narrow -- 8 bits per component, grey / grey+alpha /
 * RGB / RGBA, no palette -- so this is a small reader/writer for exactly that, CPU side, outside the hot path.
 */
#ifndef PNG_MIN_H
#define PNG_MIN_H

#include <stddef.h>
#include <stdint.h>

/* Reads 'path'. On success returns 0 and a malloc'd interleaved image (free() it). On failure returns non-zero and
 * writes a message to err. Interlaced, paletted and non-8-bit files are refused like the reference tool does. */
int png_min_read(const char* path, uint8_t** out_pixels, size_t* out_w, size_t* out_h, size_t* out_channels, char* err,
                 size_t err_len);

/* Writes an 8-bit image of 1..4 channels. effort 1..10 as in akodec's -e (more = smaller file, slower). */
int png_min_write(const char* path, const uint8_t* pixels, size_t w, size_t h, size_t channels, int effort, char* err,
                  size_t err_len);

/* Only encodes (path == NULL semantics of the tools: "all operations take place then the result is discarded"). */
int png_min_encode(const uint8_t* pixels, size_t w, size_t h, size_t channels, int effort, uint8_t** out,
                   size_t* out_size, char* err, size_t err_len);

#endif
