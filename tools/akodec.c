/*
 * akodec -- .ako -> PNG, command-line compatible with the reference's decoder tool (tools/akodec.cpp:239-343),
 * linked against libako_b200 (every stage on the GPU). PNG writing stays on the CPU (png_min.c over zlib).
 */
#include <zlib.h>

#include "ako.h"
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
	O_EFFORT,
	O_BENCHMARK,
	O_CHECKSUM,
	O_COUNT
};

static void print_version(void)
{
	printf("Ako decoding tool v%i.%i.%i\n", TOOL_MAJOR, TOOL_MINOR, TOOL_PATCH);
	printf(" - libako v%i.%i.%i, format %i (ako_b200, CUDA sm_100a)\n", akoVersionMajor(), akoVersionMinor(),
	       akoVersionPatch(), akoFormatVersion());
	printf(" - zlib %s\n", zlibVersion());
}

int main(int argc, const char* argv[])
{
	struct cli_option opt[O_COUNT] = {
	    [O_VERSION] = {"-v", "--version", CLI_FLAG, NULL, 0, 0, 0, NULL, "Print program version."},
	    [O_HELP] = {"-h", "--help", CLI_FLAG, NULL, 0, 0, 0, NULL, "Print this help."},
	    [O_VERBOSE] = {"-verbose", "--verbose", CLI_FLAG, NULL, 0, 0, 0, NULL, "Print all available information while decoding."},
	    [O_QUIET] = {"-quiet", "--quiet", CLI_FLAG, NULL, 0, 0, 0, NULL, "Don't print anything."},
	    [O_INPUT] = {"-i", "--input", CLI_TEXT, NULL, 0, 0, 0, "", "Input filename."},
	    [O_OUTPUT] = {"-o", "--output", CLI_TEXT, NULL, 0, 0, 0, "",
	                  "Output filename. If not specified, all operations will take place then the result will be discarded."},
	    [O_EFFORT] = {"-e", "--effort", CLI_INT, NULL, 1, 10, 7, NULL, "Computational effort to encode output, from 1 to 10."},
	    [O_BENCHMARK] = {"-b", "--benchmark", CLI_FLAG, NULL, 0, 0, 0, NULL, ""},
	    [O_CHECKSUM] = {"-ch", "--checksum", CLI_FLAG, NULL, 0, 0, 0, NULL, ""},
	};

	if (cli_parse(opt, O_COUNT, argc, argv) != 0)
		return 1;
	if (opt[O_HELP].value)
	{
		printf("USAGE\n");
		printf("    akodec [optional options] -i <input filename> -o <output filename>\n");
		printf("    akodec [optional options] -i <input filename>\n");
		printf("\n    Only PNG files supported as output.\n\n");
		cli_help(opt, O_COUNT);
		return 0;
	}
	if (opt[O_VERSION].value)
	{
		print_version();
		return 0;
	}

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

	uint8_t* blob = NULL;
	size_t blob_size = 0;
	{
		FILE* fp = fopen(input, "rb");
		if (fp == NULL)
		{
			printf("Error at opening file '%s'\n", input);
			return 1;
		}
		fseek(fp, 0, SEEK_END);
		const long end = ftell(fp);
		fseek(fp, 0, SEEK_SET);
		if (end < 0 || (blob = malloc((size_t)end + 1)) == NULL || fread(blob, 1, (size_t)end, fp) != (size_t)end)
		{
			printf("Error at reading file '%s'\n", input);
			fclose(fp);
			return 1;
		}
		fclose(fp);
		blob_size = (size_t)end;
	}

	struct akoSettings settings;
	size_t channels = 0, w = 0, h = 0;
	uint8_t* image = NULL;
	{
		struct stopwatch total = {{0, 0}, 0.0};
		struct stage_watches stages;
		struct akoCallbacks callbacks = akoDefaultCallbacks();
		enum akoStatus status = AKO_ERROR;
		memset(&stages, 0, sizeof(stages));
		memset(&settings, 0, sizeof(settings));
		if (benchmark && !quiet)
		{
			stopwatch_start(&total, 1);
			callbacks.events = cli_events;
			callbacks.events_data = &stages;
			printf("Benchmark: \n");
		}
		image = akoDecodeExt(&callbacks, blob_size, blob, &settings, &channels, &w, &h, &status);
		if (benchmark && !quiet)
			stopwatch_stop(&total, 1, " - Total: ");
		if (image == NULL)
		{
			printf("Ako error: '%s'\n", akoStatusString(status));
			free(blob);
			return 1;
		}
	}
	free(blob);

	if (verbose)
		printf("Input data: %zu channels, %zux%zu px, wavelet: %i, color: %i, wrap: %i, compression: %i\n", channels, w, h,
		       (int)settings.wavelet, (int)settings.color, (int)settings.wrap, (int)settings.compression);

	/* adler32 takes a 32-bit length: images of 4 GiB and more go in pieces */
	uint32_t input_checksum = 0;
	if (checksum)
	{
		uLong a = adler32(0L, Z_NULL, 0);
		for (size_t left = w * h * channels, at = 0; left > 0;)
		{
			const size_t piece = (left > ((size_t)1 << 30)) ? ((size_t)1 << 30) : left;
			a = adler32(a, image + at, (uInt)piece);
			at += piece;
			left -= piece;
		}
		input_checksum = (uint32_t)a;
	}

	if (verbose)
		printf("Encoding...\n");
	char err[128];
	uint8_t* png = NULL;
	size_t png_size = 0;
	if (png_min_encode(image, w, h, channels, opt[O_EFFORT].value, &png, &png_size, err, sizeof(err)) != 0)
	{
		printf("%s\n", err);
		return 1;
	}
	if (output[0] != '\0')
	{
		if (verbose)
			printf("Writing output: '%s'...\n", output);
		if (cli_write_blob(output, png, png_size) != 0)
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
		printf("%.2f kB <- %.2f kB, ratio: %.2f:1, %.4f bpp\n", uncompressed / 1000.0, compressed / 1000.0,
		       uncompressed / compressed, bpp);
	}

	free(png);
	akoDefaultFree(image);
	return 0;
}
