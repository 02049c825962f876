/*
 * gscan_double.c -- TEST DOUBLE of the engine ABI (include/gscan.h) on top of the CPU oracle.
 *
 * Purpose: exercise the HOST side of the product (grab_b200/host: the FileGrep mirror, its batching,
 * its scan lanes and output sequencer, the command line) on a box without a GPU.  tests/test_hostcheck.py
 * links the sources of grab_b200/host against this file instead of libgscan.so into tests/_build/grab-hostcheck and
 * compares its stdout with the bytes recorded from the unmodified reference (tests/golden/kat.json).
 *
 * This is test infrastructure, like everything under oracle/: it is never built by the product Makefile,
 * never shipped, and grab_b200/bin/grab-b200 cannot load it (that binary fails loudly without a GPU).
 * Only the entry points the host code calls are provided.  Lanes get a pseudo-random delay so that the
 * output sequencer is actually exercised (batches finish out of order).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/gscan.h"
#include "../../oracle/grab_oracle.h"

struct gscan_pattern { go_regex *re; int strict; };
struct gscan_ctx { int device; char err[256]; gscan_stats st; unsigned rng; };

static __thread char t_err[256];

int gscan_compile(const char *pattern, size_t len, uint32_t flags, gscan_pattern **out)
{
	char err[200];
	go_regex *re = go_compile(pattern, len, (flags & GSCAN_LITERAL) ? GO_LITERAL : 0u, err, sizeof err);
	if (!re) { snprintf(t_err, sizeof t_err, "%s", err); return -1; }
	if (go_nullable(re)) { snprintf(t_err, sizeof t_err, "pattern can match the empty string"); go_free(re); return -1; }
	gscan_pattern *p = calloc(1, sizeof *p);
	p->re = re;
	p->strict = (flags & GSCAN_STRICT_REF) != 0;
	*out = p;
	return 0;
}
void gscan_free_pattern(gscan_pattern *p) { if (p) { go_free(p->re); free(p); } }
int gscan_minlen(const gscan_pattern *p) { return go_minlen(p->re); }
const char *gscan_last_error(void) { return t_err; }

gscan_ctx *gscan_open(int device)
{
	const char *max = getenv("GSCAN_DOUBLE_NDEV"); /* devices the double pretends to have */
	if (device < 0 || device >= (max ? atoi(max) : 8)) { snprintf(t_err, sizeof t_err, "gscan_open: no such device %d", device); return NULL; }
	gscan_ctx *c = calloc(1, sizeof *c);
	c->device = device;
	c->rng = 12345u + (unsigned)device * 7919u;
	return c;
}
void gscan_close(gscan_ctx *ctx) { free(ctx); }
const char *gscan_why(const gscan_ctx *ctx) { return ctx->err; }

int gscan_scan_batch(gscan_ctx *ctx, const gscan_pattern *pat, const gscan_unit *units, size_t n_units, uint32_t mode,
                     gscan_match **out, size_t *n_out)
{
	go_matches m = {0};
	uint64_t bytes = 0;
	const char *fail = getenv("GSCAN_DOUBLE_FAIL_DEVICE"); /* error-path test: scans on this device fail */
	if (fail && atoi(fail) == ctx->device) { snprintf(ctx->err, sizeof ctx->err, "injected failure on device %d", ctx->device); return -1; }
	for (size_t i = 0; i < n_units; i++) {
		bytes += units[i].len;
		const uint8_t *ptr = units[i].ptr;
		uint8_t *tmp = NULL;
		if (units[i].flags & GSCAN_UNIT_FD) { /* descriptor unit: the window is read, not mapped */
			ptr = tmp = malloc(units[i].len ? units[i].len : 1);
			for (uint64_t done = 0; done < units[i].len;) {
				ssize_t r = pread((int)(intptr_t)units[i].ptr, tmp + done, units[i].len - done, (off_t)(units[i].base_off + done));
				if (r <= 0) {
					snprintf(ctx->err, sizeof ctx->err, "gscan: reading a descriptor unit: %s", r < 0 ? "error" : "file shorter than the unit");
					free(tmp);
					go_matches_free(&m);
					return -1;
				}
				done += (uint64_t)r;
			}
		}
		const int src = go_scan_window(pat->re, ptr, units[i].len, units[i].base_off, units[i].file_id, (int)mode, pat->strict, &m);
		free(tmp);
		if (src != 0) {
			snprintf(ctx->err, sizeof ctx->err, "oracle limit");
			go_matches_free(&m);
			return -1;
		}
	}
	if (getenv("GSCAN_DOUBLE_JITTER")) {
		ctx->rng = ctx->rng * 1664525u + 1013904223u;
		struct timespec ts = {0, (long)((ctx->rng >> 16) % 20000u) * 1000L}; /* 0..20 ms */
		nanosleep(&ts, NULL);
	}
	gscan_match *r = malloc((m.n ? m.n : 1) * sizeof *r);
	for (size_t i = 0; i < m.n; i++) { r[i].start = m.v[i].start; r[i].file_id = m.v[i].unit; r[i].match_len = m.v[i].len; }
	memset(&ctx->st, 0, sizeof ctx->st);
	ctx->st.bytes_scanned = bytes;
	ctx->st.n_matches = m.n;
	ctx->st.n_units = (uint32_t)n_units;
	*out = r;
	*n_out = m.n;
	go_matches_free(&m);
	return 0;
}
void gscan_free_matches(gscan_ctx *ctx, gscan_match *m) { (void)ctx; free(m); }
int gscan_last_stats(const gscan_ctx *ctx, gscan_stats *out) { *out = ctx->st; return 0; }
