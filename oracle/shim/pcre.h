/*
 * oracle/shim/pcre.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference (grab master) calls five PCRE1 functions
 *   pcre_compile      /root/reference/src/grab.cc:106
 *   pcre_study        /root/reference/src/grab.cc:115
 *   pcre_fullinfo     /root/reference/src/grab.cc:120
 *   pcre_exec         /root/reference/src/grab.cc:178
 *   pcre_free_study   /root/reference/src/grab.cc:79
 * but neither PCRE1 nor its headers exist in this image.  The image does ship
 * the PCRE2 10.42 runtime (libpcre2-8.so.0, JIT enabled) without headers.  This
 * header lets the reference's sources compile UNMODIFIED by mapping that PCRE1
 * surface onto hand-declared pcre2_*_8 prototypes.
 *
 * One ovector pair is allocated, which mirrors `int ovector[3]` / ovecsize 3 of
 * grab.cc:171,178: a pattern with a capturing group makes the match call
 * return 0 ("ovector too small"), which the reference treats as "stop"
 * (grab.cc:179) -- quirk Q2 of SURVEY.md section 8(a).
 *
 * Numeric constants are from PCRE2 10.x's public pcre2.h:
 *   PCRE2_JIT_COMPLETE   0x00000001
 *   PCRE2_INFO_MINLENGTH 16
 */
#ifndef GRAB_ORACLE_PCRE_SHIM_H
#define GRAB_ORACLE_PCRE_SHIM_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

extern "C" {
struct pcre2_real_code_8;
struct pcre2_real_match_data_8;
pcre2_real_code_8 *pcre2_compile_8(const unsigned char *, size_t, uint32_t, int *, size_t *, void *);
int pcre2_jit_compile_8(pcre2_real_code_8 *, uint32_t);
int pcre2_pattern_info_8(const pcre2_real_code_8 *, uint32_t, void *);
pcre2_real_match_data_8 *pcre2_match_data_create_8(uint32_t, void *);
void pcre2_match_data_free_8(pcre2_real_match_data_8 *);
int pcre2_match_8(const pcre2_real_code_8 *, const unsigned char *, size_t, size_t, uint32_t,
                  pcre2_real_match_data_8 *, void *);
size_t *pcre2_get_ovector_pointer_8(pcre2_real_match_data_8 *);
}

/* the two opaque PCRE1 handle types the reference stores (grab.h:52-53) */
struct real_pcre { pcre2_real_code_8 *code; };
typedef struct real_pcre pcre;
struct pcre_extra { pcre2_real_match_data_8 *mdata; };

#define PCRE_STUDY_JIT_COMPILE 0x0001
#define PCRE_INFO_MINLENGTH 15 /* PCRE1's selector value; the shim keys on it */

static inline const unsigned char *pcre_maketables(void) { return NULL; }

static inline pcre *pcre_compile(const char *pattern, int /*options: always 0*/, const char **errptr,
                                 int *erroffset, const unsigned char * /*tables*/)
{
	int ec = 0;
	size_t eo = 0;
	pcre2_real_code_8 *c =
	    pcre2_compile_8((const unsigned char *)pattern, strlen(pattern), 0, &ec, &eo, NULL);
	if (!c) {
		if (errptr) *errptr = "pcre2_compile failed";
		if (erroffset) *erroffset = (int)eo;
		return NULL;
	}
	pcre *h = new pcre;
	h->code = c;
	return h;
}

static inline pcre_extra *pcre_study(const pcre *h, int options, const char **errptr)
{
	if (options & PCRE_STUDY_JIT_COMPILE)
		pcre2_jit_compile_8(h->code, 0x00000001u);
	pcre_extra *e = new pcre_extra;
	e->mdata = pcre2_match_data_create_8(1, NULL);
	if (!e->mdata) {
		if (errptr) *errptr = "pcre2_match_data_create failed";
		delete e;
		return NULL;
	}
	return e;
}

static inline void pcre_free_study(pcre_extra *e)
{
	if (!e) return;
	pcre2_match_data_free_8(e->mdata);
	delete e;
}

static inline int pcre_fullinfo(const pcre *h, const pcre_extra *, int what, void *where)
{
	if (what != PCRE_INFO_MINLENGTH) return -3;
	uint32_t v = 0;
	int rc = pcre2_pattern_info_8(h->code, 16u, &v);
	*(int *)where = (int)v;
	return rc;
}

static inline int pcre_exec(const pcre *h, const pcre_extra *e, const char *subject, int length,
                            int startoffset, int /*options*/, int *ovector, int /*ovecsize*/)
{
	int rc = pcre2_match_8(h->code, (const unsigned char *)subject, (size_t)length,
	                       (size_t)startoffset, 0, e->mdata, NULL);
	if (rc >= 0) {
		size_t *ov = pcre2_get_ovector_pointer_8(e->mdata);
		ovector[0] = (int)ov[0];
		ovector[1] = (int)ov[1];
	}
	return rc;
}

#endif
