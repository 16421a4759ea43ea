/* png_copy in.png out.png [effort] -- reads and rewrites a PNG with png_min only. Test helper, no GPU involved. */
#include <stdio.h>
#include <stdlib.h>

#include "png_min.h"

int main(int argc, const char* argv[])
{
	if (argc < 3)
	{
		fprintf(stderr, "usage: png_copy in.png out.png [effort]\n");
		return 2;
	}
	uint8_t* px;
	size_t w, h, ch;
	char err[128];
	if (png_min_read(argv[1], &px, &w, &h, &ch, err, sizeof(err)) != 0)
	{
		fprintf(stderr, "%s\n", err);
		return 1;
	}
	if (png_min_write(argv[2], px, w, h, ch, argc > 3 ? atoi(argv[3]) : 7, err, sizeof(err)) != 0)
	{
		fprintf(stderr, "%s\n", err);
		return 1;
	}
	printf("%zu %zu %zu\n", w, h, ch);
	free(px);
	return 0;
}
