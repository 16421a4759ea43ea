/*
 * akoenc -- PNG -> .ako, command-line compatible with the reference's encoder tool (tools/akoenc.cpp:330-462:
 * same options, defaults, messages), linked against libako_b200 (every stage on the GPU). PNG reading stays on
 * the CPU (png_min.c). "-dev-r N" runs the multi-pass ratio search of EncodePass (akoenc.cpp:111-213) through
 * akoB200EncodeRatio: same blob as the reference tool, one wavelet transform instead of one per pass.
 */
#include <zlib.h>

#include "ako.h"
#include "ako_b200.h"
#include "cli.h"
#include "png_min.h"

#define TOOL_MAJOR 0
#define TOOL_MINOR 2
#define TOOL_PATCH 0

enum
{
	O_VERSION,
	O_HELP,
	O_VERBOSE,
	O_QUIET,
	O_INPUT,
	O_OUTPUT,
	O_QUANTIZATION,
	O_GATE,
	O_WAVELET,
	O_COLOR,
	O_WRAP,
	O_CHROMA_LOSS,
	O_DISCARD,
	O_BENCHMARK,
	O_CHECKSUM,
	O_RATIO,
	O_COMPRESSION,
	O_COUNT
};

static void print_version(void)
{
	printf("Ako encoding tool v%i.%i.%i\n", TOOL_MAJOR, TOOL_MINOR, TOOL_PATCH);
	printf(" - libako v%i.%i.%i, format %i (ako_b200, CUDA sm_100a)\n", akoVersionMajor(), akoVersionMinor(),
	       akoVersionPatch(), akoFormatVersion());
	printf(" - zlib %s\n", zlibVersion());
}

int main(int argc, const char* argv[])
{
	struct cli_option opt[O_COUNT] = {
	    [O_VERSION] = {"-v", "--version", CLI_FLAG, NULL, 0, 0, 0, NULL, "Print program version."},
	    [O_HELP] = {"-h", "--help", CLI_FLAG, NULL, 0, 0, 0, NULL, "Print this help."},
	    [O_VERBOSE] = {"-verbose", "--verbose", CLI_FLAG, NULL, 0, 0, 0, NULL, "Print all available information while encoding."},
	    [O_QUIET] = {"-quiet", "--quiet", CLI_FLAG, NULL, 0, 0, 0, NULL, "Don't print anything."},
	    [O_INPUT] = {"-i", "--input", CLI_TEXT, NULL, 0, 0, 0, "", "Input filename."},
	    [O_OUTPUT] = {"-o", "--output", CLI_TEXT, NULL, 0, 0, 0, "",
	                  "Output filename. If not specified, all operations will take place then the result will be discarded."},
	    [O_QUANTIZATION] = {"-q", "--quantization", CLI_INT, NULL, 0, 8192, 16, NULL,
	                        "Loss through reduced wavelet coefficient accuracy. Zero for lossless compression."},
	    [O_GATE] = {"-g", "--noise-gate", CLI_INT, NULL, 0, 8192, 0, NULL,
	                "Loss through removal of wavelet coefficients under a threshold. Zero for lossless compression."},
	    [O_WAVELET] = {"-w", "--wavelet", CLI_CHOICE, "DD137 CDF53 HAAR NONE", 0, 0, 0, NULL, "Wavelet transformation to apply."},
	    [O_COLOR] = {"-c", "--color", CLI_CHOICE, "YCOCG SUBTRACT-G NONE", 0, 0, 0, NULL, "Color transformation to apply."},
	    [O_WRAP] = {"-wr", "--wrap", CLI_CHOICE, "CLAMP MIRROR REPEAT ZERO", 0, 0, 0, NULL, "How loss wraps around image borders."},
	    [O_CHROMA_LOSS] = {"-chroma-loss", "--chroma-loss", CLI_INT, NULL, 0, 8192, 1, NULL,
	                       "Extra loss on chroma channels. Zero to disable it."},
	    [O_DISCARD] = {"-d", "--discard-non-visible", CLI_FLAG, NULL, 0, 0, 0, NULL, "Discard pixels in transparent areas."},
	    [O_BENCHMARK] = {"-b", "--benchmark", CLI_FLAG, NULL, 0, 0, 0, NULL, ""},
	    [O_CHECKSUM] = {"-ch", "--checksum", CLI_FLAG, NULL, 0, 0, 0, NULL, ""},
	    [O_RATIO] = {"-dev-r", "--dev-ratio", CLI_INT, NULL, 0, 4096, 0, NULL, ""},
	    [O_COMPRESSION] = {"-dev-compression", "--dev-compression", CLI_CHOICE, "KAGARI MANBAVARAN NONE", 0, 0, 0, NULL,
	                       "Compression method."},
	};

	if (cli_parse(opt, O_COUNT, argc, argv) != 0)
		return 1;
	if (opt[O_HELP].value)
	{
		printf("USAGE\n");
		printf("    akoenc [optional options] -i <input filename> -o <output filename>\n");
		printf("    akoenc [optional options] -i <input filename>\n");
		printf("\n    Only PNG files supported as input.\n\n");
		cli_help(opt, O_COUNT);
		return 0;
	}
	if (opt[O_VERSION].value)
	{
		print_version();
		return 0;
	}

	struct akoSettings settings = akoDefaultSettings();
	settings.quantization = opt[O_QUANTIZATION].value;
	settings.gate = opt[O_GATE].value;
	settings.discard_non_visible = opt[O_DISCARD].value;
	settings.wavelet = (enum akoWavelet)opt[O_WAVELET].value;
	settings.color = (enum akoColor)opt[O_COLOR].value;
	settings.wrap = (enum akoWrap)opt[O_WRAP].value;
	settings.chroma_loss = opt[O_CHROMA_LOSS].value;
	settings.compression = (enum akoCompression)opt[O_COMPRESSION].value;
	const int ratio = opt[O_RATIO].value;
	const int verbose = opt[O_VERBOSE].value, quiet = opt[O_QUIET].value;
	const int benchmark = opt[O_BENCHMARK].value, checksum = opt[O_CHECKSUM].value;
	const char* input = opt[O_INPUT].text;
	const char* output = opt[O_OUTPUT].text;

	if (input[0] == '\0')
	{
		printf("No input filename specified\n");
		return 1;
	}
	if (verbose)
	{
		print_version();
		printf("Opening input: '%s'...\n", input);
	}

	uint8_t* pixels = NULL;
	size_t w = 0, h = 0, channels = 0;
	char err[128];
	if (png_min_read(input, &pixels, &w, &h, &channels, err, sizeof(err)) != 0)
	{
		printf("Png error: '%s'\n", err);
		return 1;
	}
	if (verbose)
		printf("Input data: %zu channels, %zux%zu px\n", channels, w, h);

	/* adler32 takes a 32-bit length: images of 4 GiB and more go in pieces */
	uint32_t input_checksum = 0;
	if (checksum)
	{
		uLong a = adler32(0L, Z_NULL, 0);
		for (size_t left = w * h * channels, at = 0; left > 0;)
		{
			const size_t piece = (left > ((size_t)1 << 30)) ? ((size_t)1 << 30) : left;
			a = adler32(a, pixels + at, (uInt)piece);
			at += piece;
			left -= piece;
		}
		input_checksum = (uint32_t)a;
	}

	if (verbose)
	{
		printf("Encoding...\n");
		printf("[Wavelet: %i, color: %i, wrap: %i, compression %i, chroma loss: %i, discard non-visible: %i]\n",
		       (int)settings.wavelet, (int)settings.color, (int)settings.wrap, (int)settings.compression,
		       settings.chroma_loss, settings.discard_non_visible);
	}

	void* blob = NULL;
	size_t blob_size = 0;
	{
		struct stopwatch total = {{0, 0}, 0.0};
		struct stage_watches stages;
		struct akoCallbacks callbacks = akoDefaultCallbacks();
		enum akoStatus status = AKO_ERROR;
		memset(&stages, 0, sizeof(stages));

		if (benchmark && !quiet)
		{
			stopwatch_start(&total, 1);
			if (ratio == 0)
			{
				callbacks.events = cli_events;
				callbacks.events_data = &stages;
				printf("Benchmark: \n");
			}
		}

		if (ratio == 0)
			blob_size = akoEncodeExt(&callbacks, &settings, channels, w, h, pixels, &blob, &status);
		else
		{
			int used_q = 0;
			size_t passes = 0;
			if (verbose && ratio > 1)
			{
				const size_t target = (w * h * channels) / (size_t)ratio;
				printf("Target: %.2f kB, error: %.2f kB...\n", (double)target / 1000.0, (double)((target * 4) / 100) / 1000.0);
			}
			blob_size = akoB200EncodeRatio(&callbacks, &settings, ratio, channels, w, h, pixels, &blob, &used_q, &passes,
			                               &status);
			if (verbose && blob_size != 0)
				printf(" - Q: %i (%zu passes)\n", used_q, passes);
		}

		if (benchmark && !quiet)
		{
			if (ratio != 0)
				printf("Benchmark: \n");
			stopwatch_stop(&total, 1, " - Total: ");
		}
		if (blob_size == 0)
		{
			printf("Ako error: '%s'\n", akoStatusString(status));
			free(pixels);
			return 1;
		}
	}

	if (output[0] != '\0')
	{
		if (verbose)
			printf("Writing output: '%s'...\n", output);
		if (cli_write_blob(output, blob, blob_size) != 0)
		{
			printf("Write error\n");
			return 1;
		}
	}

	if (!quiet)
	{
		const double uncompressed = (double)(w * h * channels), compressed = (double)blob_size;
		const double bpp = (compressed / uncompressed) * 8.0 * (double)channels;
		if (checksum)
			printf("(%08x) ", input_checksum);
		printf("%.2f kB -> %.2f kB, ratio: %.2f:1, %.4f bpp\n", uncompressed / 1000.0, compressed / 1000.0,
		       uncompressed / compressed, bpp);
	}

	akoDefaultFree(blob);
	free(pixels);
	return 0;
}
