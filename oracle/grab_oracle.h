/*
 * grab_oracle.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle port") of the one hot
 * path of stealth/grab: the per-chunk match loop of FileGrep::find
 * (/root/reference/src/grab.cc:131-239) together with the part of libpcre (PCRE1, an external,
 * version-unpinned dependency of the reference: /root/reference/src/Makefile:14) that loop
 * relies on -- leftmost-first backtracking matching with options == 0 (grab.cc:106) and
 * PCRE_INFO_MINLENGTH (grab.cc:120).
 *
 * Parity status: PINNED against the unmodified reference built by `make -C oracle ref`
 * (grab master + oracle/shim/pcre.h -> libpcre2-8.so.0 10.42 JIT); see tests/golden/ and
 * tests/test_oracle_vs_ref.py.  The reference itself ships no tests or golden vectors
 * (SURVEY.md section 4), so the vectors were produced by running that binary here.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use this library, and only as the checker.  The product (grab_b200/) never links,
 * loads or executes anything in oracle/.
 */
#ifndef GRAB_ORACLE_H
#define GRAB_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct go_regex go_regex;

/* compile flags */
#define GO_LITERAL 1u /* treat the pattern bytes as a literal string (README.md:26, -S) */

/* pcre_compile(pattern, 0, ...) + pcre_study + MINLENGTH restated (grab.cc:101-123).
 * Returns NULL and fills err on syntax the oracle does not model. */
go_regex *go_compile(const char *pattern, size_t len, unsigned flags, char *err, size_t errlen);
void go_free(go_regex *re);
int go_minlen(const go_regex *re);        /* == PCRE_INFO_MINLENGTH for the modelled subset */
int go_capture_count(const go_regex *re); /* >0 triggers quirk Q2 in strict mode */
int go_nullable(const go_regex *re);      /* pattern can match the empty string (quirk Q4) */

/* pcre_exec(re, extra, subject, length, 0, 0, ovector, 3) restated (grab.cc:178):
 * leftmost-first match over subject[0..length).  Returns 1 and *s,*e on a match,
 * 0 on no match, -1 on internal limits. */
int go_exec(const go_regex *re, const uint8_t *subject, size_t length, size_t *s, size_t *e);

/* scan modes == which branch of grab.cc:185-212 drives `start` */
enum {
	GO_MODE_ALL = 0,   /* -O -l : every non-overlapping match                       (grab.cc:185,209) */
	GO_MODE_FIRST = 1, /* -s, or -l without -O : first match of the window only     (grab.cc:204-212) */
	GO_MODE_LINE = 2   /* line printing on: resume after the printed line remainder (grab.cc:188-209) */
};

typedef struct {
	uint64_t start; /* absolute offset: off + (start - content) + ovector[0]   (grab.cc:186) */
	uint32_t len;   /* ovector[1] - ovector[0] */
	uint32_t unit;  /* caller-supplied unit id */
} go_match;

typedef struct {
	go_match *v;
	size_t n, cap;
} go_matches;

void go_matches_free(go_matches *m);

/* One scan unit == one mmap window [base_off, base_off+clen) of grab.cc:154-169.
 * strict_q2 != 0 reproduces "capturing group => rc==0 => break" (grab.cc:179).
 * Appends to *out.  Returns 0, or -1 on error (nullable pattern: the reference would hang, Q4). */
int go_scan_window(const go_regex *re, const uint8_t *w, size_t clen, uint64_t base_off,
                   uint32_t unit, int mode, int strict_q2, go_matches *out);

typedef struct {
	int print_offset; /* -O */
	int print_line;   /* !-l */
	int single;       /* -s */
	int colored;      /* -I on a tty */
	int strict_q2;
	const char *path_prefix; /* non-NULL: "path:" is prepended to every record (grab.cc:182) */
	size_t chunk_size;       /* d_chunk_size (grab.h:48) */
} go_opts;

/* The whole of FileGrep::find for an in-memory file image: minlen skip (grab.cc:133-135),
 * chunk loop with 4096-byte overlap (:151-159), match loop, formatting (:182-207) and per-chunk
 * flush (:217-234).  Writes exactly the bytes the reference writes to stdout. */
int go_grab_buffer(const go_regex *re, const go_opts *o, const uint8_t *file, size_t size, FILE *out);

#ifdef __cplusplus
}
#endif
#endif
