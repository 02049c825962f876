/*
 * grab_oracle_main.c -- TEST INFRASTRUCTURE ONLY.  Command-line face of the oracle port with the
 * non-recursive subset of the reference's flags (/root/reference/src/main.cc:116-153), so that
 * its stdout can be diffed against oracle/_ref/grab_ref on real files.
 *   grab_oracle [-O] [-l] [-s] [-L]... [-S] [-Q] <regex> <path> [path...]
 * -S: literal pattern; -Q: do NOT reproduce quirk Q2 (capturing groups print nothing).
 */
#define _GNU_SOURCE
#include "grab_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int main(int argc, char **argv)
{
	go_opts o;
	memset(&o, 0, sizeof(o));
	o.print_line = 1;
	o.strict_q2 = 1;
	o.chunk_size = (size_t)1 << 30;                 /* main.cc:114 */
	unsigned cflags = 0;
	int c;
	while ((c = getopt(argc, argv, "OlsLSQ")) != -1) {
		switch (c) {
		case 'O': o.print_offset = 1; break;         /* main.cc:125 */
		case 'l': o.print_line = 0; break;           /* main.cc:128 */
		case 's': o.single = 1; break;               /* main.cc:122 */
		case 'L':                                    /* main.cc:131-136 */
			o.chunk_size >>= 1;
			if (o.chunk_size < ((size_t)1 << 25)) o.chunk_size = (size_t)1 << 25;
			break;
		case 'S': cflags |= GO_LITERAL; break;
		case 'Q': o.strict_q2 = 0; break;
		default: fprintf(stderr, "usage\n"); return 1;
		}
	}
	if (argc < optind + 2) { fprintf(stderr, "usage\n"); return 1; }
	const char *pat = argv[optind++];
	char err[256];
	go_regex *re = go_compile(pat, strlen(pat), cflags, err, sizeof(err));
	if (!re) { fprintf(stderr, "oracle: %s\n", err); return 255; }
	int multi = argc - optind > 1;                   /* main.cc:249-250 */
	for (; optind < argc; optind++) {
		FILE *f = fopen(argv[optind], "rb");
		if (!f) { perror(argv[optind]); return 255; }
		fseek(f, 0, SEEK_END);
		long sz = ftell(f);
		fseek(f, 0, SEEK_SET);
		uint8_t *buf = (uint8_t *)malloc(sz > 0 ? (size_t)sz : 1);
		if (sz > 0 && fread(buf, 1, (size_t)sz, f) != (size_t)sz) { perror("read"); return 255; }
		fclose(f);
		o.path_prefix = multi ? argv[optind] : NULL;
		if (go_grab_buffer(re, &o, buf, (size_t)sz, stdout) < 0) { fprintf(stderr, "oracle: scan error\n"); return 255; }
		free(buf);
	}
	go_free(re);
	return 0;
}
