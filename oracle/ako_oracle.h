/*
 * ako_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the Ako (libako v0.2.0, format 2) encode/decode hot path,
 * used as the parity checker for the CUDA implementation. Nothing in the
 * product (ako_b200/) may include, link or call this. Only tests/, bench.py's
 * cpu_baseline leg and __graft_entry__.smoke() load it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_ref.py compares every function
 * here with the unmodified reference compiled from the C files of /root/reference/library
 * (oracle/_ref/libako_ref.so, recipe in oracle/Makefile) and with the SHA-256
 * known answers of SURVEY.md Appendix B (tests/golden/kat.json).
 *
 * Layout convention of this oracle: planes are dense (plane stride = w*h, no
 * "planes spacing"); the coefficient stream is the reference's exact stream.
 */
#ifndef AKO_ORACLE_H
#define AKO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Settings mirror (plain ints so ctypes can fill it without the enum types). */
typedef struct
{
	int wavelet;     /* 0 DD137, 1 CDF53, 2 HAAR, 3 NONE   (ako.h:43-49) */
	int color;       /* 0 YCOCG, 1 SUBTRACT_G, 2 NONE, 3 YCOCG_Q (ako.h:51-58) */
	int wrap;        /* 0 CLAMP, 1 MIRROR, 2 REPEAT, 3 ZERO (ako.h:60-66) */
	int compression; /* 0 KAGARI, 1 MANBAVARAN, 2 NONE (ako.h:68-73) */
	uint64_t tiles_dimension;
	int quantization;
	int gate;
	int chroma_loss;
	int discard_non_visible;
} orc_settings;

/* SURVEY.md Appendix C synthetic image */
void orc_synth_rgba8(uint32_t w, uint32_t h, uint32_t seed, uint8_t* out);

/* geometry (misc.c:98-203) */
size_t orc_half(size_t v);
size_t orc_tile_data_size(size_t w, size_t h);
size_t orc_levels(size_t w, size_t h);
size_t orc_tile_dimension(size_t pos, size_t image_d, size_t tiles_dimension);
size_t orc_tiles_no(size_t w, size_t h, size_t tiles_dimension);

/* quantiser schedule (quantization.c:43-98) */
int16_t orc_quantization(int factor, int mul, size_t tile_w, size_t tile_h, size_t cur_w, size_t cur_h);
int16_t orc_gate(int factor, int mul, size_t tile_w, size_t tile_h, size_t cur_w, size_t cur_h);

/* format (format.c) ; planes dense, in_stride in pixels */
void orc_format_forward(int discard, int color, size_t channels, size_t w, size_t h, size_t in_stride_px,
                        const uint8_t* in, int16_t* planes);
void orc_format_inverse(int color, size_t channels, size_t w, size_t h, size_t out_stride_px, int16_t* planes,
                        uint8_t* out);

/* 1-D lifting steps, generic formulation; x has n samples, outputs t=ceil(n/2) lp and hp */
void orc_lift_1d(int wavelet, int wrap, size_t n, const int16_t* x, size_t xstride, int16_t* lp, int16_t* hp,
                 size_t ostride);
void orc_unlift_1d(int wavelet, int wrap, size_t n, const int16_t* lp, const int16_t* hp, size_t istride,
                   int16_t* x, size_t xstride);

/* multi-level 2-D (lifting.c) : dense planes <-> coefficient stream of orc_tile_data_size*channels bytes */
void orc_lift(const orc_settings* s, size_t channels, size_t w, size_t h, int16_t* planes /* destroyed */,
              int16_t* stream);
void orc_unlift(const orc_settings* s, size_t channels, size_t w, size_t h, int16_t* stream /* destroyed */,
                int16_t* planes);

/* Kagari (kagari.c) : returns bytes written, 0 on "does not fit" per the reference rule */
size_t orc_kagari_encode(size_t n_values, const int16_t* in, size_t out_cap, uint8_t* out);
/* returns bytes consumed exactly like the reference's accumulator would report, 0 on failure */
size_t orc_kagari_decode(size_t n_values, size_t in_size, const uint8_t* in, int16_t* out);
/* bit length of the Kagari stream without producing it */
uint64_t orc_kagari_bits(size_t n_values, const int16_t* in);

/* container + whole codec (encode.c / decode.c / head.c / compression.c) */
int orc_head_write(size_t channels, size_t w, size_t h, const orc_settings* s, uint8_t out[16]);
int orc_head_read(const uint8_t in[16], size_t* channels, size_t* w, size_t* h, orc_settings* s);
/* returns blob size (0 on error, status in *status). out must hold orc_encode_bound() bytes */
size_t orc_encode_bound(size_t channels, size_t w, size_t h);
size_t orc_encode(const orc_settings* s, size_t channels, size_t w, size_t h, const uint8_t* in, uint8_t* out,
                  int* status);
/* out must hold w*h*channels bytes (query dims with orc_head_read first); returns status */
int orc_decode(size_t in_size, const uint8_t* in, uint8_t* out, orc_settings* out_s);

/* multi-pass ratio search of the encoder tool (tools/akoenc.cpp:111-213, EncodePass) */
size_t orc_encode_pass(int ratio, const orc_settings* s, size_t channels, size_t w, size_t h, const uint8_t* in,
                       uint8_t* out, int* out_q, size_t* out_passes, int* status);

#ifdef __cplusplus
}
#endif
#endif
