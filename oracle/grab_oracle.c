/*
 * grab_oracle.c -- TEST INFRASTRUCTURE ONLY (see grab_oracle.h).
 *
 * Part 1 restates the slice of PCRE the reference leans on (external libpcre, linked by
 * /root/reference/src/Makefile:14 and called at grab.cc:106,115,120,178): a byte-oriented,
 * options==0 (no CASELESS/MULTILINE/DOTALL/UTF, "C"-locale tables from pcre_maketables(),
 * grab.cc:106) Perl-compatible backtracking matcher with leftmost-first semantics, plus the
 * study-time minimum-length figure.  The published algorithm restated here is the classic
 * one of PCRE's interpreter: try each start offset left to right; inside, alternatives in
 * source order, greedy quantifiers take as much as possible and give back one item at a
 * time.  It is written as a small backtracking VM with an explicit stack so that very long
 * greedy runs do not recurse.
 *
 * Part 2 restates FileGrep::find (grab.cc:131-239) line by line.
 */
#define _GNU_SOURCE
#include "grab_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* byte sets                                                                                   */
/* ------------------------------------------------------------------------------------------ */

typedef struct { uint32_t w[8]; } bset;

static void bs_clear(bset *s) { memset(s, 0, sizeof(*s)); }
static void bs_add(bset *s, unsigned c) { s->w[(c & 255) >> 5] |= 1u << (c & 31); }
static int bs_has(const bset *s, unsigned c) { return (s->w[(c & 255) >> 5] >> (c & 31)) & 1u; }
static void bs_range(bset *s, unsigned lo, unsigned hi) { for (unsigned c = lo; c <= hi; c++) bs_add(s, c); }
static void bs_or(bset *d, const bset *a) { for (int i = 0; i < 8; i++) d->w[i] |= a->w[i]; }
static void bs_not(bset *s) { for (int i = 0; i < 8; i++) s->w[i] = ~s->w[i]; }
static void bs_fill(bset *s) { memset(s, 0xff, sizeof(*s)); }
static int bs_empty(const bset *s) { for (int i = 0; i < 8; i++) if (s->w[i]) return 0; return 1; }

/* "C"-locale ctype classes as pcre_maketables() builds them without setlocale() */
static int c_isupper(unsigned c) { return c >= 'A' && c <= 'Z'; }
static int c_islower(unsigned c) { return c >= 'a' && c <= 'z'; }
static int c_isalpha(unsigned c) { return c_isupper(c) || c_islower(c); }
static int c_isdigit(unsigned c) { return c >= '0' && c <= '9'; }
static int c_isalnum(unsigned c) { return c_isalpha(c) || c_isdigit(c); }
static int c_isword(unsigned c) { return c_isalnum(c) || c == '_'; }
/* \s : PCRE >= 8.34 and PCRE2 include VT (0x0b) */
static int c_isspace(unsigned c) { return c == ' ' || (c >= 9 && c <= 13); }
static int c_isxdigit(unsigned c) { return c_isdigit(c) || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'); }
static int c_ispunct(unsigned c) { return c > 32 && c < 127 && !c_isalnum(c); }
static int c_isprint(unsigned c) { return c >= 32 && c < 127; }
static int c_isgraph(unsigned c) { return c > 32 && c < 127; }
static int c_iscntrl(unsigned c) { return c < 32 || c == 127; }
static int c_isblank(unsigned c) { return c == ' ' || c == '\t'; }

static void bs_pred(bset *s, int (*p)(unsigned), int negate)
{
	for (unsigned c = 0; c < 256; c++)
		if ((p(c) != 0) != (negate != 0)) bs_add(s, c);
}

static void bs_caseless(bset *s)
{
	for (unsigned c = 'a'; c <= 'z'; c++) {
		if (bs_has(s, c)) bs_add(s, c - 32);
		if (bs_has(s, c - 32)) bs_add(s, c);
	}
}

/* ------------------------------------------------------------------------------------------ */
/* AST                                                                                         */
/* ------------------------------------------------------------------------------------------ */

enum { N_EMPTY, N_SET, N_CAT, N_ALT, N_REP, N_GROUP, N_ASSERT, N_LOOK /* (?=) (?!) (?<=) (?<!) */, N_ATOMIC /* (?>...) and possessive groups */ };
enum { A_BOL, A_EOL, A_SOS, A_EOS, A_EOSNL, A_WORDB, A_NWORDB, A_MBOL, A_MEOL };
enum { Q_GREEDY, Q_LAZY, Q_POSSESSIVE };

typedef struct node {
	int type;
	bset set;            /* N_SET */
	struct node **kid;   /* N_CAT / N_ALT children; N_REP / N_GROUP use kid[0] */
	int nkid, capkid;
	uint32_t rmin, rmax; /* N_REP; rmax == UINT32_MAX: unbounded */
	int qkind;           /* N_REP */
	int capturing;       /* N_GROUP */
	int akind;           /* N_ASSERT */
	int ahead, neg;      /* N_LOOK */
} node;

#define INF UINT32_MAX

typedef struct {
	const uint8_t *p, *end;
	char *err;
	size_t errlen;
	int failed;
	int ncapture;
	int depth;
} parser;

static void perr(parser *P, const char *msg)
{
	if (!P->failed && P->err && P->errlen) snprintf(P->err, P->errlen, "%s", msg);
	P->failed = 1;
}

static node *nnew(int type)
{
	node *n = (node *)calloc(1, sizeof(node));
	n->type = type;
	return n;
}

static void nadd(node *n, node *k)
{
	if (n->nkid == n->capkid) {
		n->capkid = n->capkid ? n->capkid * 2 : 4;
		n->kid = (node **)realloc(n->kid, sizeof(node *) * (size_t)n->capkid);
	}
	n->kid[n->nkid++] = k;
}

static void nfree(node *n)
{
	if (!n) return;
	for (int i = 0; i < n->nkid; i++) nfree(n->kid[i]);
	free(n->kid);
	free(n);
}

typedef struct { int icase, dotall, multiline, ungreedy, extended; } pflags;

static node *parse_alt(parser *P, pflags f);

static int hexval(int c)
{
	if (c >= '0' && c <= '9') return c - '0';
	if (c >= 'a' && c <= 'f') return c - 'a' + 10;
	if (c >= 'A' && c <= 'F') return c - 'A' + 10;
	return -1;
}

/* Parses the escape after a backslash.  Returns 1 and a single byte in *ch, 2 and a set in
 * *set (class escape), 3 for an assertion in *ch, 0 on error. `incls`: inside [...] */
static int parse_escape(parser *P, int incls, unsigned *ch, bset *set)
{
	if (P->p >= P->end) { perr(P, "\\ at end of pattern"); return 0; }
	unsigned c = *P->p++;
	bs_clear(set);
	switch (c) {
	case 'd': bs_pred(set, c_isdigit, 0); return 2;
	case 'D': bs_pred(set, c_isdigit, 1); return 2;
	case 'w': bs_pred(set, c_isword, 0); return 2;
	case 'W': bs_pred(set, c_isword, 1); return 2;
	case 's': bs_pred(set, c_isspace, 0); return 2;
	case 'S': bs_pred(set, c_isspace, 1); return 2;
	case 'h': bs_add(set, ' '); bs_add(set, '\t'); bs_add(set, 0xa0); return 2;
	case 'H': bs_add(set, ' '); bs_add(set, '\t'); bs_add(set, 0xa0); bs_not(set); return 2;
	case 'v': bs_range(set, 10, 13); bs_add(set, 0x85); return 2;
	case 'V': bs_range(set, 10, 13); bs_add(set, 0x85); bs_not(set); return 2;
	case 'N':
		if (incls) { perr(P, "\\N in class"); return 0; }
		bs_fill(set); set->w[0] &= ~(1u << 10); return 2;
	case 'n': *ch = '\n'; return 1;
	case 't': *ch = '\t'; return 1;
	case 'r': *ch = '\r'; return 1;
	case 'f': *ch = '\f'; return 1;
	case 'a': *ch = 7; return 1;
	case 'e': *ch = 27; return 1;
	case 'c':
		if (P->p >= P->end) { perr(P, "\\c at end"); return 0; }
		{
			unsigned x = *P->p++;
			if (x >= 'a' && x <= 'z') x -= 32;
			*ch = x ^ 0x40;
		}
		return 1;
	case 'x': {
		unsigned v = 0;
		if (P->p < P->end && *P->p == '{') {
			const uint8_t *q = P->p + 1;
			int nd = 0;
			while (q < P->end && hexval(*q) >= 0) { v = v * 16 + (unsigned)hexval(*q); q++; nd++; }
			if (q >= P->end || *q != '}' || nd == 0 || v > 255) { perr(P, "bad \\x{..}"); return 0; }
			P->p = q + 1;
		} else {
			int nd = 0;
			while (nd < 2 && P->p < P->end && hexval(*P->p) >= 0) { v = v * 16 + (unsigned)hexval(*P->p); P->p++; nd++; }
		}
		*ch = v;
		return 1;
	}
	case '0': {
		unsigned v = 0;
		int nd = 0;
		while (nd < 2 && P->p < P->end && *P->p >= '0' && *P->p <= '7') { v = v * 8 + (unsigned)(*P->p - '0'); P->p++; nd++; }
		*ch = v & 255;
		return 1;
	}
	case 'b':
		if (incls) { *ch = 8; return 1; }
		*ch = A_WORDB; return 3;
	case 'B': if (incls) { perr(P, "\\B in class"); return 0; } *ch = A_NWORDB; return 3;
	case 'A': if (incls) { perr(P, "\\A in class"); return 0; } *ch = A_SOS; return 3;
	case 'z': if (incls) { perr(P, "\\z in class"); return 0; } *ch = A_EOS; return 3;
	case 'Z': if (incls) { perr(P, "\\Z in class"); return 0; } *ch = A_EOSNL; return 3;
	default:
		if (c >= '1' && c <= '9') { perr(P, "back references are not modelled"); return 0; }
		if (c_isalnum(c)) { perr(P, "escape sequence not modelled"); return 0; }
		*ch = c; /* any escaped non-alphanumeric is itself */
		return 1;
	}
}

static const struct { const char *name; int (*pred)(unsigned); } posix_classes[] = {
	{"alpha", c_isalpha}, {"digit", c_isdigit}, {"alnum", c_isalnum}, {"upper", c_isupper},
	{"lower", c_islower}, {"space", c_isspace}, {"xdigit", c_isxdigit}, {"punct", c_ispunct},
	{"print", c_isprint}, {"graph", c_isgraph}, {"cntrl", c_iscntrl}, {"blank", c_isblank},
	{"word", c_isword},
};

static node *parse_class(parser *P, pflags f)
{
	/* P->p is just past '[' */
	node *n = nnew(N_SET);
	int negate = 0, first = 1;
	if (P->p < P->end && *P->p == '^') { negate = 1; P->p++; }
	for (;;) {
		if (P->p >= P->end) { perr(P, "missing ] in class"); return n; }
		unsigned c = *P->p;
		if (c == ']' && !first) { P->p++; break; }
		first = 0;
		unsigned lo;
		int have_lo = 0;
		if (c == '[' && P->p + 1 < P->end && P->p[1] == ':') {
			const uint8_t *q = P->p + 2;
			int neg = 0;
			if (q < P->end && *q == '^') { neg = 1; q++; }
			const uint8_t *name = q;
			while (q < P->end && *q >= 'a' && *q <= 'z') q++;
			if (q + 1 < P->end && q[0] == ':' && q[1] == ']') {
				size_t nl = (size_t)(q - name);
				int found = 0;
				for (size_t i = 0; i < sizeof(posix_classes) / sizeof(posix_classes[0]); i++)
					if (strlen(posix_classes[i].name) == nl && !memcmp(posix_classes[i].name, name, nl)) {
						bset t; bs_clear(&t); bs_pred(&t, posix_classes[i].pred, neg); bs_or(&n->set, &t);
						found = 1;
					}
				if (!found) { perr(P, "unknown POSIX class"); return n; }
				P->p = q + 2;
				continue;
			}
			/* not a POSIX class: '[' is a literal */
		}
		if (c == '\\') {
			P->p++;
			bset t;
			unsigned ch;
			int k = parse_escape(P, 1, &ch, &t);
			if (k == 0) return n;
			if (k == 2) { bs_or(&n->set, &t); continue; }
			lo = ch; have_lo = 1;
		} else {
			lo = c; have_lo = 1; P->p++;
		}
		if (have_lo) {
			/* range? */
			if (P->p + 1 < P->end && P->p[0] == '-' && P->p[1] != ']') {
				const uint8_t *save = P->p;
				P->p++;
				unsigned hi;
				if (*P->p == '\\') {
					P->p++;
					bset t;
					int k = parse_escape(P, 1, &hi, &t);
					if (k == 0) return n;
					if (k == 2) { /* "a-\d": '-' is literal */
						bs_add(&n->set, lo); bs_add(&n->set, '-'); bs_or(&n->set, &t);
						continue;
					}
				} else if (*P->p == '[' && P->p + 1 < P->end && P->p[1] == ':') {
					P->p = save; bs_add(&n->set, lo); continue;
				} else {
					hi = *P->p++;
				}
				if (hi < lo) { perr(P, "range out of order in class"); return n; }
				bs_range(&n->set, lo, hi);
			} else {
				bs_add(&n->set, lo);
			}
		}
	}
	if (f.icase) bs_caseless(&n->set);
	if (negate) bs_not(&n->set);
	return n;
}

static node *mkset_char(unsigned c, pflags f)
{
	node *n = nnew(N_SET);
	bs_add(&n->set, c);
	if (f.icase) bs_caseless(&n->set);
	return n;
}

static node *mkassert(int k)
{
	node *n = nnew(N_ASSERT);
	n->akind = k;
	return n;
}

/* try to read {n}, {n,}, {n,m} at P->p (which points at '{'); returns 1 if it is a quantifier */
static int parse_braces(parser *P, uint32_t *mn, uint32_t *mx)
{
	const uint8_t *q = P->p + 1;
	if (q >= P->end || !c_isdigit(*q)) return 0;
	unsigned long a = 0, b = 0;
	while (q < P->end && c_isdigit(*q)) { a = a * 10 + (unsigned long)(*q - '0'); if (a > 65535) return -1; q++; }
	if (q < P->end && *q == '}') { *mn = *mx = (uint32_t)a; P->p = q + 1; return 1; }
	if (q >= P->end || *q != ',') return 0;
	q++;
	if (q < P->end && *q == '}') { *mn = (uint32_t)a; *mx = INF; P->p = q + 1; return 1; }
	if (q >= P->end || !c_isdigit(*q)) return 0;
	while (q < P->end && c_isdigit(*q)) { b = b * 10 + (unsigned long)(*q - '0'); if (b > 65535) return -1; q++; }
	if (q >= P->end || *q != '}') return 0;
	if (b < a) return -1;
	*mn = (uint32_t)a; *mx = (uint32_t)b; P->p = q + 1;
	return 1;
}

static node *parse_atom(parser *P, pflags *f, int *is_flag_change)
{
	*is_flag_change = 0;
	unsigned c = *P->p++;
	switch (c) {
	case '(': {
		int capturing = 1;
		pflags inner = *f;
		if (P->p < P->end && *P->p == '?') {
			P->p++;
			if (P->p >= P->end) { perr(P, "bad (?"); return NULL; }
			if (*P->p == ':') { capturing = 0; P->p++; }
			else if (*P->p == 'P' && P->p + 1 < P->end && P->p[1] == '<') {
				P->p += 2;
				while (P->p < P->end && *P->p != '>') P->p++;
				if (P->p >= P->end) { perr(P, "bad named group"); return NULL; }
				P->p++;
			} else if ((*P->p == '<' || *P->p == '\'') && P->p + 1 < P->end && P->p[1] != '=' && P->p[1] != '!') {
				unsigned close = *P->p == '<' ? '>' : '\'';
				P->p++;
				while (P->p < P->end && *P->p != close) P->p++;
				if (P->p >= P->end) { perr(P, "bad named group"); return NULL; }
				P->p++;
			} else if (*P->p == '#') { /* (?#comment): up to the next ')' */
				while (P->p < P->end && *P->p != ')') P->p++;
				if (P->p >= P->end) { perr(P, "missing ) after comment"); return NULL; }
				P->p++;
				*is_flag_change = 1;
				return NULL;
			} else if (*P->p == '=' || *P->p == '!' || *P->p == '>' ||
			           (*P->p == '<' && P->p + 1 < P->end && (P->p[1] == '=' || P->p[1] == '!'))) {
				/* lookahead (?= (?!, lookbehind (?<= (?<!, atomic group (?> : none of them captures */
				int kind = 0; /* 0 ahead, 1 behind, 2 atomic */
				int neg = 0;
				if (*P->p == '>') { kind = 2; P->p++; }
				else if (*P->p == '<') { kind = 1; neg = P->p[1] == '!'; P->p += 2; }
				else { neg = *P->p == '!'; P->p++; }
				if (++P->depth > 200) { perr(P, "nesting too deep"); return NULL; }
				node *body = parse_alt(P, inner);
				P->depth--;
				if (P->failed) { nfree(body); return NULL; }
				if (P->p >= P->end || *P->p != ')') { perr(P, "missing )"); nfree(body); return NULL; }
				P->p++;
				node *g = nnew(kind == 2 ? N_ATOMIC : N_LOOK);
				g->ahead = kind == 0;
				g->neg = neg;
				nadd(g, body);
				return g;
			} else if (*P->p == '<' || *P->p == '|' ||
			           *P->p == 'R' || *P->p == '(' || c_isdigit(*P->p) || *P->p == '&' ||
			           *P->p == 'C' || *P->p == '+') {
				perr(P, "group construct not modelled (recursion/conditional)");
				return NULL;
			} else {
				/* inline options: (?i) (?-i) (?is-m) (?i:...) */
				int on = 1;
				pflags nf = *f;
				for (;;) {
					if (P->p >= P->end) { perr(P, "bad inline option"); return NULL; }
					unsigned o = *P->p++;
					if (o == '-') { on = 0; continue; }
					if (o == 'i') { nf.icase = on; continue; }
					if (o == 's') { nf.dotall = on; continue; }
					if (o == 'm') { nf.multiline = on; continue; }
					if (o == 'U') { nf.ungreedy = on; continue; } /* PCRE_UNGREEDY: greedy <-> lazy */
					if (o == 'x') { nf.extended = on; continue; } /* PCRE_EXTENDED: white space and #-comments ignored */
					if (o == ')') { *f = nf; *is_flag_change = 1; return NULL; }
					if (o == ':') { inner = nf; capturing = 0; break; }
					perr(P, "inline option not modelled");
					return NULL;
				}
			}
		}
		if (capturing) P->ncapture++;
		if (++P->depth > 200) { perr(P, "nesting too deep"); return NULL; }
		node *body = parse_alt(P, inner);
		P->depth--;
		if (P->failed) { nfree(body); return NULL; }
		if (P->p >= P->end || *P->p != ')') { perr(P, "missing )"); nfree(body); return NULL; }
		P->p++;
		node *g = nnew(N_GROUP);
		g->capturing = capturing;
		nadd(g, body);
		return g;
	}
	case '[': return parse_class(P, *f);
	case '.': {
		node *n = nnew(N_SET);
		bs_fill(&n->set);
		if (!f->dotall) n->set.w[0] &= ~(1u << 10);
		return n;
	}
	case '^': return mkassert(f->multiline ? A_MBOL : A_BOL);
	case '$': return mkassert(f->multiline ? A_MEOL : A_EOL);
	case '\\': {
		if (P->p < P->end && *P->p == 'Q') {
			/* \Q...\E literal run */
			P->p++;
			node *cat = nnew(N_CAT);
			while (P->p < P->end) {
				if (P->p + 1 < P->end && P->p[0] == '\\' && P->p[1] == 'E') { P->p += 2; break; }
				nadd(cat, mkset_char(*P->p++, *f));
			}
			return cat;
		}
		if (P->p < P->end && *P->p == 'E') { P->p++; return nnew(N_EMPTY); }
		bset t;
		unsigned ch;
		int k = parse_escape(P, 0, &ch, &t);
		if (k == 0) return NULL;
		if (k == 1) return mkset_char(ch, *f);
		if (k == 2) { node *n = nnew(N_SET); n->set = t; return n; }
		return mkassert((int)ch);
	}
	case '*': case '+': case '?':
		perr(P, "quantifier does not follow a repeatable item");
		return NULL;
	default:
		return mkset_char(c, *f);
	}
}

/* (?x): outside character classes white space is ignored and # starts a comment that ends at the next newline */
/* (?#...) comments vanish wherever they stand, also between an item and its quantifier ("a(?#x)+" is "a+") */
static void skip_extended(parser *P, const pflags *f)
{
	while (P->p < P->end) {
		unsigned c = *P->p;
		if (c == '(' && P->p + 2 < P->end && P->p[1] == '?' && P->p[2] == '#') {
			const uint8_t *q = P->p + 3;
			while (q < P->end && *q != ')') q++;
			if (q >= P->end) return; /* unterminated: parse_atom reports it */
			P->p = q + 1;
			continue;
		}
		if (!f->extended) break;
		if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v') { P->p++; continue; }
		if (c == '#') { while (P->p < P->end && *P->p != '\n') P->p++; continue; }
		break;
	}
}

static node *parse_concat(parser *P, pflags *f)
{
	node *cat = nnew(N_CAT);
	for (;;) {
		skip_extended(P, f);
		if (P->failed || P->p >= P->end || *P->p == '|' || *P->p == ')') break;
		int flagchange = 0;
		node *a = parse_atom(P, f, &flagchange);
		if (P->failed) { nfree(a); break; }
		if (flagchange) continue;
		if (!a) break;
		/* quantifier */
		for (;;) {
			skip_extended(P, f);
			if (P->p >= P->end) break;
			uint32_t mn = 0, mx = 0;
			unsigned q = *P->p;
			int isq = 0;
			if (q == '*') { mn = 0; mx = INF; P->p++; isq = 1; }
			else if (q == '+') { mn = 1; mx = INF; P->p++; isq = 1; }
			else if (q == '?') { mn = 0; mx = 1; P->p++; isq = 1; }
			else if (q == '{') {
				int r = parse_braces(P, &mn, &mx);
				if (r < 0) { perr(P, "bad {n,m} quantifier"); nfree(a); a = NULL; break; }
				isq = r;
			}
			if (!isq) break;
			int kind = f->ungreedy ? Q_LAZY : Q_GREEDY;
			if (P->p < P->end && *P->p == '?') { kind = f->ungreedy ? Q_GREEDY : Q_LAZY; P->p++; }
			else if (P->p < P->end && *P->p == '+') { kind = Q_POSSESSIVE; P->p++; }
			if (a->type == N_ASSERT || a->type == N_LOOK) { perr(P, "quantified assertion not modelled"); nfree(a); a = NULL; break; }
			node *r = nnew(N_REP);
			r->rmin = mn; r->rmax = mx; r->qkind = kind;
			nadd(r, a);
			a = r;
			{ /* X*+ on anything but a single byte class is the atomic group (?>X*) */
				const node *in = r->kid[0];
				while (in->type == N_GROUP && !in->capturing) in = in->kid[0];
				while ((in->type == N_CAT || in->type == N_ALT) && in->nkid == 1) in = in->kid[0];
				if (kind == Q_POSSESSIVE && in->type != N_SET) {
					r->qkind = Q_GREEDY;
					node *at = nnew(N_ATOMIC);
					nadd(at, r);
					a = at;
				}
			}
			break; /* "a{2}{3}" style stacking is not modelled: next loop iteration would see it as literal/err */
		}
		if (!a) break;
		nadd(cat, a);
	}
	return cat;
}

static node *parse_alt(parser *P, pflags f)
{
	node *alt = nnew(N_ALT);
	pflags cur = f;
	for (;;) {
		node *c = parse_concat(P, &cur);
		nadd(alt, c);
		if (P->failed) break;
		if (P->p < P->end && *P->p == '|') { P->p++; continue; } /* option changes persist across | (PCRE) */
		break;
	}
	if (alt->nkid == 1) {
		node *only = alt->kid[0];
		alt->nkid = 0;
		nfree(alt);
		return only;
	}
	return alt;
}

/* ------------------------------------------------------------------------------------------ */
/* AST properties                                                                              */
/* ------------------------------------------------------------------------------------------ */

static uint64_t n_minlen(const node *n)
{
	uint64_t m, t;
	switch (n->type) {
	case N_EMPTY: case N_ASSERT: return 0;
	case N_SET: return 1;
	case N_CAT: m = 0; for (int i = 0; i < n->nkid; i++) m += n_minlen(n->kid[i]); return m;
	case N_ALT:
		m = UINT64_MAX;
		for (int i = 0; i < n->nkid; i++) { t = n_minlen(n->kid[i]); if (t < m) m = t; }
		return m;
	case N_REP: return (uint64_t)n->rmin * n_minlen(n->kid[0]);
	case N_GROUP: case N_ATOMIC: return n_minlen(n->kid[0]);
	case N_LOOK: return 0;
	}
	return 0;
}

static int n_nullable(const node *n) { return n_minlen(n) == 0; }

/* 1 and *len if every match of n has the same length (what a lookbehind branch must have) */
static int n_fixedlen(const node *n, uint64_t *len)
{
	uint64_t a, b;
	switch (n->type) {
	case N_EMPTY: case N_ASSERT: case N_LOOK: *len = 0; return 1;
	case N_SET: *len = 1; return 1;
	case N_CAT:
		a = 0;
		for (int i = 0; i < n->nkid; i++) { if (!n_fixedlen(n->kid[i], &b)) return 0; a += b; }
		*len = a;
		return 1;
	case N_ALT:
		if (!n_fixedlen(n->kid[0], &a)) return 0;
		for (int i = 1; i < n->nkid; i++) if (!n_fixedlen(n->kid[i], &b) || b != a) return 0;
		*len = a;
		return 1;
	case N_REP:
		if (n->rmin != n->rmax || !n_fixedlen(n->kid[0], &a)) return 0;
		*len = a * n->rmin;
		return 1;
	case N_GROUP: case N_ATOMIC: return n_fixedlen(n->kid[0], len);
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* program                                                                                     */
/* ------------------------------------------------------------------------------------------ */

enum { I_SET, I_SPLIT, I_JMP, I_REP, I_ASSERT, I_MATCH, I_CAP /* a capturing group closed */,
       I_LOOK /* a: LK_*, b: bytes to step back (lookbehind), c: pc behind the construct */, I_LOOKEND };
enum { LK_AHEAD, LK_AHEAD_NEG, LK_BEHIND, LK_BEHIND_NEG, LK_ATOMIC };

typedef struct { uint32_t op, a, b, c, d; } inst;

struct go_regex {
	inst *prog;
	size_t nprog, capprog;
	bset *sets;
	size_t nsets, capsets;
	int minlen, ncapture, nullable;
	bset first;       /* bytes a match can start with (valid iff first_valid) */
	int first_valid;
	int failed;
	char *err;
	size_t errlen;
};

#define MAXPROG (1u << 20)

static uint32_t emit(go_regex *re, uint32_t op, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
	if (re->nprog >= MAXPROG) {
		if (!re->failed && re->err) snprintf(re->err, re->errlen, "pattern too large");
		re->failed = 1;
		return 0;
	}
	if (re->nprog == re->capprog) {
		re->capprog = re->capprog ? re->capprog * 2 : 64;
		re->prog = (inst *)realloc(re->prog, re->capprog * sizeof(inst));
	}
	inst *i = &re->prog[re->nprog];
	i->op = op; i->a = a; i->b = b; i->c = c; i->d = d;
	return (uint32_t)re->nprog++;
}

static uint32_t addset(go_regex *re, const bset *s)
{
	for (size_t i = 0; i < re->nsets; i++)
		if (!memcmp(&re->sets[i], s, sizeof(bset))) return (uint32_t)i;
	if (re->nsets == re->capsets) {
		re->capsets = re->capsets ? re->capsets * 2 : 16;
		re->sets = (bset *)realloc(re->sets, re->capsets * sizeof(bset));
	}
	re->sets[re->nsets] = *s;
	return (uint32_t)re->nsets++;
}

static void gen(go_regex *re, const node *n);

static void gen_rep(go_regex *re, const node *n)
{
	const node *k = n->kid[0];
	/* strip non-capturing wrappers around a single set */
	const node *inner = k;
	while (inner->type == N_GROUP && !inner->capturing) inner = inner->kid[0];
	while ((inner->type == N_CAT || inner->type == N_ALT) && inner->nkid == 1) inner = inner->kid[0];
	if (inner->type == N_SET) {
		emit(re, I_REP, addset(re, &inner->set), n->rmin, n->rmax, (uint32_t)n->qkind);
		return;
	}
	if (n->qkind == Q_POSSESSIVE) {
		if (!re->failed && re->err) snprintf(re->err, re->errlen, "possessive group quantifier not modelled");
		re->failed = 1;
		return;
	}
	if (n->rmax == INF && n_nullable(k)) {
		if (!re->failed && re->err) snprintf(re->err, re->errlen, "unbounded repeat of an empty-matchable group not modelled");
		re->failed = 1;
		return;
	}
	for (uint32_t i = 0; i < n->rmin && !re->failed; i++) gen(re, k);
	if (n->rmax == INF) {
		/* L: split body,out ; body ; jmp L */
		uint32_t L = emit(re, I_SPLIT, 0, 0, 0, 0);
		gen(re, k);
		emit(re, I_JMP, L, 0, 0, 0);
		uint32_t out = (uint32_t)re->nprog;
		if (re->failed) return;
		if (n->qkind == Q_GREEDY) { re->prog[L].a = L + 1; re->prog[L].b = out; }
		else { re->prog[L].a = out; re->prog[L].b = L + 1; }
	} else {
		uint32_t opt = n->rmax - n->rmin;
		uint32_t *splits = (uint32_t *)malloc(sizeof(uint32_t) * (opt ? opt : 1));
		for (uint32_t i = 0; i < opt && !re->failed; i++) {
			splits[i] = emit(re, I_SPLIT, 0, 0, 0, 0);
			gen(re, k);
		}
		uint32_t out = (uint32_t)re->nprog;
		for (uint32_t i = 0; i < opt && !re->failed; i++) {
			if (n->qkind == Q_GREEDY) { re->prog[splits[i]].a = splits[i] + 1; re->prog[splits[i]].b = out; }
			else { re->prog[splits[i]].a = out; re->prog[splits[i]].b = splits[i] + 1; }
		}
		free(splits);
	}
}

static void gen(go_regex *re, const node *n)
{
	if (re->failed) return;
	switch (n->type) {
	case N_EMPTY: break;
	case N_SET: emit(re, I_SET, addset(re, &n->set), 0, 0, 0); break;
	case N_ASSERT: emit(re, I_ASSERT, (uint32_t)n->akind, 0, 0, 0); break;
	case N_CAT: for (int i = 0; i < n->nkid; i++) gen(re, n->kid[i]); break;
	case N_GROUP:
		gen(re, n->kid[0]);
		if (n->capturing) emit(re, I_CAP, 0, 0, 0, 0);
		break;
	case N_ATOMIC: {
		uint32_t l = emit(re, I_LOOK, LK_ATOMIC, 0, 0, 0);
		gen(re, n->kid[0]);
		emit(re, I_LOOKEND, 0, 0, 0, 0);
		if (!re->failed) re->prog[l].c = (uint32_t)re->nprog;
		break;
	}
	case N_LOOK: {
		if (n->ahead) {
			uint32_t l = emit(re, I_LOOK, n->neg ? LK_AHEAD_NEG : LK_AHEAD, 0, 0, 0);
			gen(re, n->kid[0]);
			emit(re, I_LOOKEND, 0, 0, 0, 0);
			if (!re->failed) re->prog[l].c = (uint32_t)re->nprog;
			break;
		}
		/* lookbehind: every top-level branch has its own fixed length (PCRE's rule); (?<=a|bc) is (?:(?<=a)|(?<=bc)),
		 * (?<!a|bc) is (?<!a)(?<!bc) */
		const node *body = n->kid[0];
		while (body->type == N_GROUP && !body->capturing) body = body->kid[0];
		int nb = body->type == N_ALT ? body->nkid : 1;
		uint32_t *jmps = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nb);
		for (int i = 0; i < nb && !re->failed; i++) {
			const node *br = body->type == N_ALT ? body->kid[i] : body;
			uint64_t len = 0;
			if (!n_fixedlen(br, &len) || len > 65535) {
				if (!re->failed && re->err) snprintf(re->err, re->errlen, "lookbehind assertion is not fixed length");
				re->failed = 1;
				break;
			}
			uint32_t split = 0;
			if (!n->neg && i + 1 < nb) split = emit(re, I_SPLIT, 0, 0, 0, 0);
			uint32_t l = emit(re, I_LOOK, n->neg ? LK_BEHIND_NEG : LK_BEHIND, (uint32_t)len, 0, 0);
			gen(re, br);
			emit(re, I_LOOKEND, 0, 0, 0, 0);
			if (re->failed) break;
			re->prog[l].c = (uint32_t)re->nprog;
			if (!n->neg && i + 1 < nb) {
				jmps[i] = emit(re, I_JMP, 0, 0, 0, 0);
				re->prog[split].a = split + 1;
				re->prog[split].b = (uint32_t)re->nprog;
			}
		}
		if (!n->neg) for (int i = 0; i + 1 < nb && !re->failed; i++) re->prog[jmps[i]].a = (uint32_t)re->nprog;
		free(jmps);
		break;
	}
	case N_REP: gen_rep(re, n); break;
	case N_ALT: {
		/* split a1, L2 ; a1 ; jmp end ; L2: split a2, L3 ; ... ; an */
		uint32_t *jmps = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n->nkid);
		for (int i = 0; i < n->nkid && !re->failed; i++) {
			if (i + 1 < n->nkid) {
				uint32_t s = emit(re, I_SPLIT, 0, 0, 0, 0);
				gen(re, n->kid[i]);
				jmps[i] = emit(re, I_JMP, 0, 0, 0, 0);
				if (re->failed) break;
				re->prog[s].a = s + 1;
				re->prog[s].b = (uint32_t)re->nprog;
			} else {
				gen(re, n->kid[i]);
			}
		}
		for (int i = 0; i + 1 < n->nkid && !re->failed; i++) re->prog[jmps[i]].a = (uint32_t)re->nprog;
		free(jmps);
		break;
	}
	}
}

/* first-byte set: bytes at which an attempt can possibly succeed.  Returns nullable-ness. */
static int n_first(const node *n, bset *out, int *unknown)
{
	switch (n->type) {
	case N_EMPTY: return 1;
	case N_ASSERT: case N_LOOK: *unknown = 1; return 1; /* keep it simple: assertions disable the skip table */
	case N_ATOMIC: return n_first(n->kid[0], out, unknown);
	case N_SET: bs_or(out, &n->set); return 0;
	case N_GROUP: return n_first(n->kid[0], out, unknown);
	case N_CAT:
		for (int i = 0; i < n->nkid; i++)
			if (!n_first(n->kid[i], out, unknown)) return 0;
		return 1;
	case N_ALT: {
		int nul = 0;
		for (int i = 0; i < n->nkid; i++) nul |= n_first(n->kid[i], out, unknown);
		return nul;
	}
	case N_REP: {
		int nul = n_first(n->kid[0], out, unknown);
		return nul || n->rmin == 0;
	}
	}
	return 1;
}

go_regex *go_compile(const char *pattern, size_t len, unsigned flags, char *err, size_t errlen)
{
	parser P;
	memset(&P, 0, sizeof(P));
	P.p = (const uint8_t *)pattern;
	P.end = P.p + len;
	P.err = err;
	P.errlen = errlen;
	if (err && errlen) err[0] = 0;
	node *root;
	pflags f0 = {0, 0, 0, 0, 0};
	if (flags & GO_LITERAL) {
		root = nnew(N_CAT);
		for (size_t i = 0; i < len; i++) nadd(root, mkset_char((uint8_t)pattern[i], f0));
	} else {
		root = parse_alt(&P, f0);
		if (!P.failed && P.p < P.end) perr(&P, "unmatched )");
	}
	if (P.failed) { nfree(root); return NULL; }

	go_regex *re = (go_regex *)calloc(1, sizeof(go_regex));
	re->err = err;
	re->errlen = errlen;
	gen(re, root);
	emit(re, I_MATCH, 0, 0, 0, 0);
	uint64_t ml = n_minlen(root);
	re->minlen = ml > 65535 ? 65535 : (int)ml; /* PCRE2 study caps the figure at 65535 */
	re->nullable = ml == 0;
	re->ncapture = P.ncapture;
	bs_clear(&re->first);
	int unknown = 0;
	int nul = n_first(root, &re->first, &unknown);
	re->first_valid = !nul && !unknown && !bs_empty(&re->first);
	nfree(root);
	if (re->failed) { go_free(re); return NULL; }
	re->err = NULL;
	return re;
}

void go_free(go_regex *re)
{
	if (!re) return;
	free(re->prog);
	free(re->sets);
	free(re);
}

int go_minlen(const go_regex *re) { return re->minlen; }
int go_capture_count(const go_regex *re) { return re->ncapture; }
int go_nullable(const go_regex *re) { return re->nullable; }

/* ------------------------------------------------------------------------------------------ */
/* backtracking VM                                                                             */
/* ------------------------------------------------------------------------------------------ */

typedef struct { uint32_t pc; uint32_t kind; size_t sp, lo; int cap; /* was a capturing group set when this choice was pushed? */ } bt;
enum { BT_PLAIN, BT_REP_GIVEBACK, BT_REP_TAKEMORE, BT_LOOK /* lo: LK_* kind, sp: position to return to, pc: behind the construct */ };

typedef struct { bt *v; size_t n, cap; } btstack;

static int bt_push(btstack *s, uint32_t pc, uint32_t kind, size_t sp, size_t lo, int cap)
{
	if (s->n == s->cap) {
		size_t nc = s->cap ? s->cap * 2 : 256;
		if (nc > ((size_t)1 << 26)) return -1;
		bt *nv = (bt *)realloc(s->v, nc * sizeof(bt));
		if (!nv) return -1;
		s->v = nv;
		s->cap = nc;
	}
	s->v[s->n].pc = pc; s->v[s->n].kind = kind; s->v[s->n].sp = sp; s->v[s->n].lo = lo; s->v[s->n].cap = cap;
	s->n++;
	return 0;
}

static int is_word_at(const uint8_t *s, size_t len, size_t i) { return i < len && c_isword(s[i]); }

static int check_assert(int kind, const uint8_t *s, size_t len, size_t sp)
{
	switch (kind) {
	case A_BOL: case A_SOS: return sp == 0;
	case A_MBOL: return sp == 0 || s[sp - 1] == '\n';
	case A_EOS: return sp == len;
	case A_EOL: case A_EOSNL: return sp == len || (sp + 1 == len && s[sp] == '\n');
	case A_MEOL: return sp == len || s[sp] == '\n';
	case A_WORDB: case A_NWORDB: {
		int before = sp > 0 && is_word_at(s, len, sp - 1);
		int after = is_word_at(s, len, sp);
		return (before != after) == (kind == A_WORDB);
	}
	}
	return 0;
}

/* one anchored attempt at offset `at`; returns 1 and *e on success (2: a capturing group was set on the successful
 * path -- with room for a single offset pair pcre_exec then returns 0, quirk Q2), 0 on failure, -1 on limit */
/* PCRE bounds its search (match limit, 10 million by default); so does the oracle, or a pathological pattern
 * would keep a test busy for hours.  Hitting the limit is reported (-1), never turned into an answer. */
#define GO_MATCH_LIMIT 20000000ul

static int attempt(const go_regex *re, const uint8_t *s, size_t len, size_t at, size_t *e, btstack *st)
{
	uint32_t pc = 0;
	size_t sp = at;
	unsigned long steps = 0;
	int cap = 0;
	st->n = 0;
	for (;;) {
		if (++steps > GO_MATCH_LIMIT) return -1;
		const inst *in = &re->prog[pc];
		int fail = 0;
		switch (in->op) {
		case I_MATCH: *e = sp; return cap ? 2 : 1;
		case I_CAP: cap = 1; pc++; break;
		case I_SET:
			if (sp < len && bs_has(&re->sets[in->a], s[sp])) { sp++; pc++; }
			else fail = 1;
			break;
		case I_ASSERT:
			if (check_assert((int)in->a, s, len, sp)) pc++;
			else fail = 1;
			break;
		case I_JMP: pc = in->a; break;
		case I_LOOK: {
			size_t back = in->b;
			if ((in->a == LK_BEHIND || in->a == LK_BEHIND_NEG) && sp < back) {
				/* fewer bytes before sp than the branch needs (the subject begins at the moving search start): it cannot match */
				if (in->a == LK_BEHIND) fail = 1; else pc = in->c;
				break;
			}
			if (bt_push(st, in->c, BT_LOOK, sp, in->a, cap) < 0) return -1;
			sp -= (in->a == LK_BEHIND || in->a == LK_BEHIND_NEG) ? back : 0;
			pc++;
			break;
		}
		case I_LOOKEND: {
			/* the body matched: the choice points it left are dropped (an assertion / atomic group is never re-entered) */
			size_t L = st->n;
			while (L > 0 && st->v[L - 1].kind != BT_LOOK) L--;
			if (L == 0) return -1; /* cannot happen: every LOOKEND has its frame */
			bt fr = st->v[L - 1];
			st->n = L - 1;
			if (fr.lo == LK_AHEAD_NEG || fr.lo == LK_BEHIND_NEG) { cap = fr.cap; fail = 1; break; } /* negative assertion is false */
			if (fr.lo != LK_ATOMIC) sp = fr.sp;                  /* assertions consume nothing */
			pc++;
			break;
		}
		case I_SPLIT:
			if (bt_push(st, in->b, BT_PLAIN, sp, 0, cap) < 0) return -1;
			pc = in->a;
			break;
		case I_REP: {
			const bset *set = &re->sets[in->a];
			size_t mn = in->b, mx = in->c == INF ? (size_t)-1 : (size_t)in->c;
			size_t avail = len - sp, k = 0;
			if (in->d == Q_LAZY) {
				while (k < mn && k < avail && bs_has(set, s[sp + k])) k++;
				if (k < mn) { fail = 1; break; }
				/* lo field carries how many more may be taken */
				if (mx > mn && bt_push(st, pc, BT_REP_TAKEMORE, sp + k, mx == (size_t)-1 ? (size_t)-1 : mx - mn, cap) < 0) return -1;
				sp += k; pc++;
			} else {
				size_t lim = mx < avail ? mx : avail;
				while (k < lim && bs_has(set, s[sp + k])) k++;
				if (k < mn) { fail = 1; break; }
				if (in->d == Q_GREEDY && k > mn && bt_push(st, pc + 1, BT_REP_GIVEBACK, sp + k, sp + mn, cap) < 0) return -1;
				sp += k; pc++;
			}
			break;
		}
		}
		if (!fail) continue;
		/* backtrack */
		for (;;) {
			if (st->n == 0) return 0;
			bt *t = &st->v[st->n - 1];
			cap = t->cap; /* groups closed after this choice point are undone */
			if (t->kind == BT_LOOK) {
				/* the body of an assertion / atomic group failed for good */
				int negative = t->lo == LK_AHEAD_NEG || t->lo == LK_BEHIND_NEG;
				uint32_t cont = t->pc;
				size_t back_to = t->sp;
				st->n--;
				if (!negative) continue;          /* positive assertion / atomic group: the failure goes on */
				pc = cont; sp = back_to;           /* negative assertion holds: go on behind it */
				break;
			}
			if (t->kind == BT_PLAIN) { pc = t->pc; sp = t->sp; st->n--; break; }
			if (t->kind == BT_REP_GIVEBACK) {
				/* t->sp: current end of the greedy run; give back one item */
				if (t->sp > t->lo) {
					t->sp--;
					pc = t->pc; sp = t->sp;
					if (t->sp == t->lo) st->n--;
					break;
				}
				st->n--;
				continue;
			}
			/* BT_REP_TAKEMORE: lazy repeat takes one more item */
			{
				const inst *ri = &re->prog[t->pc];
				if (t->lo > 0 && t->sp < len && bs_has(&re->sets[ri->a], s[t->sp])) {
					t->sp++;
					if (t->lo != (size_t)-1) t->lo--;
					pc = t->pc + 1; sp = t->sp;
					if (t->lo == 0) st->n--;
					break;
				}
				st->n--;
				continue;
			}
		}
	}
}

int go_exec(const go_regex *re, const uint8_t *subject, size_t length, size_t *s, size_t *e)
{
	btstack st = {0, 0, 0};
	int rc = 0;
	size_t at = 0;
	size_t ml = (size_t)re->minlen;
	for (; at <= length; at++) {
		if (re->first_valid) {
			while (at < length && !bs_has(&re->first, subject[at])) at++;
			if (at >= length) break;
		}
		if (length - at < ml) break; /* study's minimum-length shortcut: cannot match any more */
		size_t end = 0;
		rc = attempt(re, subject, length, at, &end, &st);
		if (rc < 0) break;
		if (rc >= 1) { *s = at; *e = end; break; }
	}
	free(st.v);
	return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Part 2: FileGrep::find restated                                                             */
/* ------------------------------------------------------------------------------------------ */

void go_matches_free(go_matches *m)
{
	free(m->v);
	m->v = NULL;
	m->n = m->cap = 0;
}

static int push_match(go_matches *m, uint64_t start, uint32_t len, uint32_t unit)
{
	if (m->n == m->cap) {
		size_t nc = m->cap ? m->cap * 2 : 64;
		go_match *nv = (go_match *)realloc(m->v, nc * sizeof(go_match));
		if (!nv) return -1;
		m->v = nv;
		m->cap = nc;
	}
	m->v[m->n].start = start;
	m->v[m->n].len = len;
	m->v[m->n].unit = unit;
	m->n++;
	return 0;
}

int go_scan_window(const go_regex *re, const uint8_t *w, size_t clen, uint64_t base_off,
                   uint32_t unit, int mode, int strict_q2, go_matches *out)
{
	if (re->nullable) return -1; /* grab.cc:209 would spin forever on an empty match (Q4) */
	size_t start = 0;
	const size_t minlen = (size_t)re->minlen;
	while (start + minlen < clen) {                         /* grab.cc:175  (strict <  => Q1) */
		size_t s = 0, e = 0;
		int rc = go_exec(re, w + start, clen - start, &s, &e); /* grab.cc:178 */
		if (rc < 0) return -1;                                /* oracle's own match limit: report, do not guess */
		if (rc == 2) rc = strict_q2 ? 0 : 1;                  /* a group was set: ovecsize 3 too small => pcre_exec returns 0 (Q2) */
		if (rc <= 0) break;                                   /* grab.cc:179 */
		if (push_match(out, base_off + start + s, (uint32_t)(e - s), unit) < 0) return -1;
		size_t a = 0;
		if (mode == GO_MODE_LINE) {                           /* grab.cc:194-196 */
			size_t p = start + e;
			while (p < clen && w[p] != '\n' && a < 511) { a++; p++; }
		} else if (mode == GO_MODE_FIRST) {
			break;                                            /* grab.cc:206 / :211 */
		}
		start += e + a;                                       /* grab.cc:209 */
	}
	return 0;
}

int go_grab_buffer(const go_regex *re, const go_opts *o, const uint8_t *file, size_t size, FILE *out)
{
	static const char start_inv[] = "\33[7m", stop_inv[] = "\33[27m"; /* grab.cc:66-67 */
	if (re->nullable) return -1;
	const size_t minlen = (size_t)re->minlen;
	if (minlen > size) return 0;                              /* grab.cc:133-135 */
	const size_t overlap = 0x1000;                            /* grab.cc:151 */
	if (o->chunk_size <= overlap) return -1;
	for (size_t off = 0; off < size; off += o->chunk_size - overlap) { /* grab.cc:154 */
		size_t clen = size - off < o->chunk_size ? size - off : o->chunk_size; /* :156-159 */
		const uint8_t *w = file + off;
		size_t start = 0;
		int printed = 0;
		while (start + minlen < clen) {                       /* grab.cc:175 */
			size_t s = 0, e = 0;
			int rc = go_exec(re, w + start, clen - start, &s, &e);
			if (rc < 0) return -1;
			if (rc == 2) rc = o->strict_q2 ? 0 : 1;
			if (rc <= 0) break;
			if (o->path_prefix) { fputs(o->path_prefix, out); fputc(':', out); } /* :182-183 */
			if (o->print_offset)                              /* :185-186 */
				fprintf(out, "Match at offset %llu\n", (unsigned long long)(off + start + s));
			size_t a = 0;
			if (o->print_line) {                              /* :189-203 */
				size_t ms = start + s, me = start + e, b = 0;
				while (ms - b > start && w[ms - b - 1] != '\n' && b < 511) b++;
				size_t p = me;
				while (p < clen && w[p] != '\n' && a < 511) { a++; p++; }
				fwrite(w + ms - b, 1, b, out);
				if (o->colored) fputs(start_inv, out);
				fwrite(w + ms, 1, me - ms, out);
				if (o->colored) fputs(stop_inv, out);
				fwrite(w + me, 1, a, out);
				fputc('\n', out);
				printed = 1;
			} else if (!o->print_offset) {                    /* :204-207 */
				fputs("matches\n", out);
				printed = 1;
				break;
			}
			printed = 1;
			start += e + a;                                   /* :209 */
			if (o->single) break;                             /* :211-212 */
		}
		/* per-chunk flush (:217-234): nothing to model but the -s early exit */
		if (printed && o->single) break;                      /* :232-233 */
	}
	return 0;
}
