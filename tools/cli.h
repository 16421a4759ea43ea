/*
 * cli.h -- what the two tools share: a table-driven option parser accepting the command lines of the reference's
 * tools (tools/options.hpp: "-q 16", "--quantization 16", flags without value, same error messages), a stopwatch
 * fed by the library's events callback (tools/benchmark.hpp:73-90) and the Adler-32 printed by -ch
 * (tools/misc.hpp:60-84; zlib's adler32 is the same function).
 */
#ifndef AKO_CLI_H
#define AKO_CLI_H

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

#include "ako.h"

enum cli_kind
{
	CLI_FLAG,
	CLI_INT,
	CLI_TEXT,  /* free text */
	CLI_CHOICE /* one of a space separated list, case-insensitive; stored as its index */
};

struct cli_option
{
	const char* short_name;
	const char* long_name;
	enum cli_kind kind;
	const char* choices; /* CLI_CHOICE */
	int min, max;        /* CLI_INT */
	int value;           /* flag / int / choice index */
	const char* text;    /* CLI_TEXT */
	const char* help;
};

static struct cli_option* cli_find(struct cli_option* table, size_t n, const char* arg)
{
	for (size_t i = 0; i < n; i++)
		if (strcmp(arg, table[i].short_name) == 0 || strcmp(arg, table[i].long_name) == 0)
			return &table[i];
	return NULL;
}

/* 0: ok */
static int cli_parse(struct cli_option* table, size_t n, int argc, const char* argv[])
{
	for (int i = 1; i < argc;)
	{
		struct cli_option* o = cli_find(table, n, argv[i]);
		if (o == NULL)
		{
			fprintf(stderr, "Error, unknown option '%s'.\n", argv[i]);
			return 1;
		}
		if (o->kind == CLI_FLAG)
		{
			o->value = 1;
			i += 1;
			continue;
		}
		if (i + 1 >= argc)
		{
			fprintf(stderr, "Error, no value specified for option '%s'.\n", argv[i]);
			return 1;
		}
		const char* v = argv[i + 1];
		int ok = 1;
		if (o->kind == CLI_INT)
		{
			char* end = NULL;
			const long x = strtol(v, &end, 10);
			ok = (end != v && *end == '\0' && x >= o->min && x <= o->max);
			if (ok)
				o->value = (int)x;
		}
		else if (o->kind == CLI_TEXT)
			o->text = v;
		else
		{
			ok = 0;
			int index = 0;
			for (const char* c = o->choices; *c != '\0'; index++)
			{
				const size_t len = strcspn(c, " ");
				if (strlen(v) == len && strncasecmp(c, v, len) == 0)
				{
					o->value = index;
					ok = 1;
					break;
				}
				c += len;
				c += strspn(c, " ");
			}
		}
		if (!ok)
		{
			fprintf(stderr, "Error, invalid value '%s' for option '%s'.\n", v, argv[i]);
			return 1;
		}
		i += 2;
	}
	return 0;
}

static void cli_help(const struct cli_option* table, size_t n)
{
	for (size_t i = 0; i < n; i++)
	{
		printf("%s, %s\n", table[i].short_name, table[i].long_name);
		if (table[i].help[0] != '\0')
			printf("    %s\n", table[i].help);
		if (table[i].kind == CLI_INT)
			printf("    Default value: %i\n", table[i].value);
		else if (table[i].kind == CLI_CHOICE)
			printf("    Options: %s\n", table[i].choices);
		printf("\n");
	}
}

/* ---- stopwatches driven by the library's events (benchmark.hpp:38-90) ---- */

struct stopwatch
{
	struct timespec last;
	double ms;
};

static void stopwatch_start(struct stopwatch* s, int fresh)
{
	clock_gettime(CLOCK_MONOTONIC, &s->last);
	if (fresh)
		s->ms = 0.0;
}

static void stopwatch_stop(struct stopwatch* s, int print, const char* name)
{
	struct timespec now;
	clock_gettime(CLOCK_MONOTONIC, &now);
	s->ms += (double)(now.tv_sec - s->last.tv_sec) * 1e3 + (double)(now.tv_nsec - s->last.tv_nsec) / 1e6;
	if (print)
		printf("%s%g ms\n", name, s->ms);
}

struct stage_watches
{
	struct stopwatch format, wavelet, compression;
};

static void cli_events(size_t tile_no, size_t total_tiles, enum akoEvent e, void* raw)
{
	struct stage_watches* sw = raw;
	const int first = (tile_no == 0), last = (tile_no == total_tiles - 1);
	switch (e)
	{
	case AKO_EVENT_FORMAT_START: stopwatch_start(&sw->format, first); break;
	case AKO_EVENT_WAVELET_START: stopwatch_start(&sw->wavelet, first); break;
	case AKO_EVENT_COMPRESSION_START: stopwatch_start(&sw->compression, first); break;
	case AKO_EVENT_FORMAT_END: stopwatch_stop(&sw->format, last, " - Format: "); break;
	case AKO_EVENT_WAVELET_END: stopwatch_stop(&sw->wavelet, last, " - Wavelet transformation: "); break;
	case AKO_EVENT_COMPRESSION_END: stopwatch_stop(&sw->compression, last, " - Compression: "); break;
	default: break;
	}
}

static int cli_write_blob(const char* path, const void* blob, size_t size)
{
	FILE* fp = fopen(path, "wb");
	if (fp == NULL)
		return 1;
	const int bad = (fwrite(blob, 1, size, fp) != size);
	return (fclose(fp) != 0) || bad;
}

#endif
