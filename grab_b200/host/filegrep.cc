// filegrep.cc -- see filegrep.h.  Host C++ only; the scan itself is gscan_scan_batch().
#include "filegrep.h"

#include <fcntl.h>
#include <ftw.h>
#include <pthread.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <mutex>
#include <system_error>
#include <thread>

#include "../../include/gscan.h"

namespace grab_b200 {

// one lock around stdout, like stdout_lock of grab.cc:56,219-225
static pthread_mutex_t g_stdout_lock = PTHREAD_MUTEX_INITIALIZER;

// GRAB_B200_TRACE=1: milestones with milliseconds since the process was loaded, on stderr (stdout is the parity surface)
static const std::chrono::steady_clock::time_point g_t0 = std::chrono::steady_clock::now();
static const bool g_trace = getenv("GRAB_B200_TRACE") != nullptr;
static void trace_ts(const char *what, int lane = -1)
{
	if (!g_trace) return;
	const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_t0).count();
	if (lane >= 0) fprintf(stderr, "[grab-b200] %9.2f ms  lane %d: %s\n", ms, lane, what);
	else fprintf(stderr, "[grab-b200] %9.2f ms  %s\n", ms, what);
}

static const char kStartInv[] = "\33[7m", kStopInv[] = "\33[27m"; // grab.cc:66-67

// One batch of queued windows on its way through a lane.
struct FileGrep::Batch {
	uint64_t seq = 0;
	std::vector<Window> windows;
};

// Job queue + output sequencer shared by the walking thread and the lanes of one FileGrep.
struct FileGrep::Pipeline {
	std::mutex mu;
	std::condition_variable cv_job, cv_space, cv_turn, cv_idle;
	std::deque<Batch *> jobs;
	uint64_t next_seq = 0, next_out = 0; // next batch number to submit / to print
	size_t inflight = 0;                 // submitted and not yet printed
	bool stop = false, failed = false;
	std::string err;
	// -s: once a window of a file has printed, the rest of that file prints nothing (grab.cc:232-233); owned by
	// whoever holds the output turn
	bool have_done = false;
	uint32_t done_seq = 0;
	std::vector<std::thread> lanes;
};

FileGrep::FileGrep() : FileGrep(0) {}

FileGrep::FileGrep(int device) : d_device(device) { d_my_uid = geteuid(); }

FileGrep::~FileGrep()
{
	flush();
	if (d_pipe) {
		{
			std::lock_guard<std::mutex> g(d_pipe->mu);
			d_pipe->stop = true;
		}
		d_pipe->cv_job.notify_all();
		for (auto &t : d_pipe->lanes) t.join();
	}
	if (d_pat) gscan_free_pattern(d_pat);
}

void FileGrep::config(const std::map<std::string, size_t> &config)
{
	if (config.count("color") > 0) d_colored = true;
	if (config.count("noline") > 0) d_print_line = false;
	if (config.count("offsets") > 0) d_print_offset = true;
	if (config.count("single") > 0) d_single_match = true;
	if (config.count("low_mem") > 0) d_low_mem = true;
	if (config.count("literal") > 0) d_literal = true;
	if (config.count("lenient") > 0) d_strict = false;
	auto it = config.find("chunk_size");
	if (it != config.end()) d_chunk_size = it->second;
	it = config.find("device");
	if (it != config.end()) d_device = (int)it->second;
	it = config.find("batch_bytes");
	if (it != config.end() && it->second >= 1) d_batch_bytes = it->second;
	it = config.find("ndev");
	if (it != config.end() && it->second >= 1) d_ndev = (int)it->second;
	it = config.find("lanes");
	if (it != config.end() && it->second >= 1) d_lanes_per_dev = (int)it->second;
	it = config.find("fd_budget");
	if (it != config.end()) d_fd_budget = it->second;
}

int FileGrep::prepare(const std::string &regex)
{
	uint32_t flags = (d_literal ? GSCAN_LITERAL : 0u) | (d_strict ? GSCAN_STRICT_REF : 0u);
	if (d_pat) { gscan_free_pattern(d_pat); d_pat = nullptr; }
	if (gscan_compile(regex.c_str(), regex.size(), flags, &d_pat) != 0) {
		// the reference reports "FileGrep::prepare::pcre_compile error" (grab.cc:107); keep that prefix, add the reason
		d_err = std::string("FileGrep::prepare::pcre_compile error (") + gscan_last_error() + ")";
		return -1;
	}
	d_minlen = gscan_minlen(d_pat);
	trace_ts("pattern compiled");
	return 0;
}

void FileGrep::release(Window &w)
{
	if (w.map) munmap(w.map, w.clen); // grab.cc:215
	w.map = nullptr;
	if (w.fd >= 0) close(w.fd);
	w.fd = -1;
}

int FileGrep::find(const char *path, const struct stat *st, int)
{
	if (!d_pat) { d_err = "FileGrep::find: no pattern prepared"; return -1; }
	size_t clen = (size_t)st->st_size;
	if ((size_t)d_minlen > clen) return 0; // grab.cc:133-135

	int fd = -1, flags = O_RDONLY | O_NOCTTY;
#ifdef __linux__
	if (st->st_uid == d_my_uid || d_my_uid == 0) flags |= O_NOATIME; // grab.cc:139-143
#endif
	fd = open(path, flags);
	if (fd < 0 && (errno == EMFILE || errno == ENFILE) && (!d_queue.empty() || d_pipe)) {
		flush_quietly(); // queued descriptor windows are what holds the descriptors: print them and try again
		fd = open(path, flags);
	}
	if (fd < 0) {
		d_err = "FileGrep::find::open: " + std::string(strerror(errno));
		return -1;
	}
	const off_t overlap = 0x1000; // grab.cc:151
	const uint32_t seq = d_file_seq++;
	// Without line output nothing on the host ever looks at the window's bytes: the window is queued as a descriptor
	// and the engine's staging threads pread() it straight into their pinned buffers (GSCAN_UNIT_FD) -- no mapping, no
	// page faults on it, no munmap.  With line output (the bytes around a match are printed) the window is mapped
	// like the reference's (grab.cc:161).
	const bool by_fd = !d_print_line && d_fd_budget >= 3 * 64;
	for (off_t off = 0; off < st->st_size; off += ((off_t)d_chunk_size - overlap)) { // grab.cc:154
		// -s: the reference stops reading a file once one of its windows printed (grab.cc:232-233).  The sequencer
		// suppresses later windows anyway; not queueing them only saves the work when an earlier batch is already out
		if (d_single_match && d_pipe) {
			std::lock_guard<std::mutex> g(d_pipe->mu);
			if (d_pipe->have_done && d_pipe->done_seq == seq) break;
		}
		clen = (st->st_size - off < (off_t)d_chunk_size) ? (size_t)(st->st_size - off) : d_chunk_size;
		Window w;
		if (by_fd) {
			const bool last = off + ((off_t)d_chunk_size - overlap) >= st->st_size;
			w.fd = last ? fd : dup(fd); // the file's last window takes the descriptor along
			if (w.fd < 0 && (errno == EMFILE || errno == ENFILE) && (!d_queue.empty() || d_pipe)) {
				flush_quietly(); // queued windows hold descriptors: print them, which closes theirs, and try again
				w.fd = dup(fd);
			}
			if (w.fd < 0) {
				d_err = "FileGrep::find::dup: " + std::string(strerror(errno));
				close(fd);
				return -1;
			}
			if (last) fd = -1;
#ifdef POSIX_FADV_WILLNEED
			if (clen > 4 * 0x1000) posix_fadvise(w.fd, off, (off_t)clen, POSIX_FADV_WILLNEED); // start the read-ahead now (a no-op on tmpfs)
#endif
		} else {
			void *m = mmap(nullptr, clen, PROT_READ, MAP_PRIVATE | MAP_NORESERVE, fd, off); // grab.cc:126-128,161
			if (m == MAP_FAILED && errno == ENOMEM && (!d_queue.empty() || d_pipe)) {
				// out of address space or mappings (vm.max_map_count) because queued windows are still mapped: the reference
				// holds one window at a time -- print what is queued, which unmaps it, and try again
				flush_quietly();
				m = mmap(nullptr, clen, PROT_READ, MAP_PRIVATE | MAP_NORESERVE, fd, off);
			}
			if (m == MAP_FAILED) {
				d_err = "FileGrep::find::mmap: " + std::string(strerror(errno));
				close(fd);
				return -1;
			}
			if (clen > 4 * 0x1000 && !d_single_match) posix_madvise(m, clen, POSIX_MADV_SEQUENTIAL); // grab.cc:168-169
			w.map = static_cast<uint8_t *>(m);
		}
		w.path = path;
		w.clen = clen;
		w.off = (uint64_t)off;
		w.file_seq = seq;
		d_queue.push_back(std::move(w));
		d_queued_bytes += clen;
		// a batch is cut by bytes or by windows: every queued window is a live mapping (or an open descriptor), and a tree of
		// tiny files would otherwise run into vm.max_map_count (65530) long before 256 MiB are queued
		if ((d_queued_bytes >= d_batch_bytes || d_queue.size() >= max_windows_per_batch()) && submit() < 0) {
			if (fd >= 0) close(fd);
			return -1;
		}
	}
	if (fd >= 0) close(fd);
	return 0;
}

// the bytes the reference appends to its ostringstream for one window (grab.cc:175-213), given the
// matches the engine selected for it
struct gscan_match_view { uint64_t start; uint32_t len; };

void FileGrep::format_window(const Window &w, const gscan_match_view *m, size_t n, std::string &out) const
{
	char num[32];
	uint64_t search_start = w.off; // `start` of grab.cc:172, as an absolute offset
	for (size_t i = 0; i < n; i++) {
		if (d_recursive || d_print_path) { out += w.path; out += ':'; }     // grab.cc:182-183
		if (d_print_offset) {                                               // grab.cc:185-186
			out += "Match at offset ";
			snprintf(num, sizeof(num), "%llu", (unsigned long long)m[i].start);
			out += num;
			out += '\n';
		}
		if (d_print_line) {                                                 // grab.cc:188-203
			const uint8_t *base = w.map;
			const size_t ms = (size_t)(m[i].start - w.off), me = ms + m[i].len, lo = (size_t)(search_start - w.off);
			size_t b = 0, a = 0;
			while (ms - b > lo && base[ms - b - 1] != '\n' && b < 511) b++;
			while (me + a < w.clen && base[me + a] != '\n' && a < 511) a++;
			out.append(reinterpret_cast<const char *>(base + ms - b), b);
			if (d_colored) out += kStartInv;
			out.append(reinterpret_cast<const char *>(base + ms), m[i].len);
			if (d_colored) out += kStopInv;
			out.append(reinterpret_cast<const char *>(base + me), a);
			out += '\n';
			search_start = w.off + me + a;                                  // grab.cc:209
		} else if (!d_print_offset) {                                       // grab.cc:204-207
			out += "matches\n";
			break;
		} else {
			search_start = m[i].start + m[i].len;
		}
		if (d_single_match) break;                                          // grab.cc:211-212
	}
}

// Moves the queued windows into a batch for the lanes.  Returns -1 only when no lane could be started; the failure of
// an earlier batch is kept for flush() to report.  The queue is empty afterwards either way.
int FileGrep::submit()
{
	if (d_queue.empty()) return 0;
	if (!d_pipe) d_pipe.reset(new Pipeline);
	Pipeline &p = *d_pipe;
	Batch *b = new Batch;
	b->windows.swap(d_queue);
	d_queued_bytes = 0;
	std::unique_lock<std::mutex> lk(p.mu);
	if (p.lanes.empty()) {
		trace_ts("first batch submitted");
		const int n = d_ndev * d_lanes_per_dev;
		for (int i = 0; i < n; i++) {
			try {
				p.lanes.emplace_back(&FileGrep::lane_main, this, i);
			} catch (const std::system_error &e) { // thread limit, no memory: work with the lanes that did start
				if (p.lanes.empty()) {
					d_err = std::string("FileGrep::find::lane: ") + e.what();
					lk.unlock();
					for (auto &w : b->windows) release(w);
					delete b;
					return -1;
				}
				break;
			}
		}
	}
	// bounded look-ahead: at most one waiting batch per lane keeps the mapped-but-unscanned windows small
	p.cv_space.wait(lk, [&] { return p.jobs.size() < p.lanes.size(); });
	// a batch that failed earlier is reported by flush(); later batches are still scanned and printed -- the reference
	// keeps going after a file it could not search (grab.cc:267-268)
	b->seq = p.next_seq++;
	p.inflight++;
	p.jobs.push_back(b);
	lk.unlock();
	p.cv_job.notify_one();
	return 0;
}

// windows per batch: all batches that can be alive at once (one being filled, one waiting and one running per lane)
// stay well below the default vm.max_map_count of 65530
size_t FileGrep::max_windows_per_batch() const
{
	const size_t lanes = (size_t)(d_ndev * d_lanes_per_dev);
	size_t cap = 49152 / (2 * lanes + 1);
	if (!d_print_line && d_fd_budget >= 3 * 64 && d_fd_budget / (2 * lanes + 1) < cap) cap = d_fd_budget / (2 * lanes + 1); // descriptor windows
	return cap < 8192 ? (cap < 64 ? 64 : cap) : 8192;
}

// flush() that keeps a lane failure for the final flush() to report
void FileGrep::flush_quietly()
{
	submit();
	if (!d_pipe) return;
	std::unique_lock<std::mutex> lk(d_pipe->mu);
	d_pipe->cv_idle.wait(lk, [&] { return d_pipe->inflight == 0; });
}

// Submits what is queued and waits until every batch has been printed.
int FileGrep::flush()
{
	int rc = submit();
	if (!d_pipe) return rc;
	Pipeline &p = *d_pipe;
	std::unique_lock<std::mutex> lk(p.mu);
	p.cv_idle.wait(lk, [&] { return p.inflight == 0; });
	trace_ts("all batches printed");
	if (p.failed) {
		d_err = p.err;
		p.failed = false; // reported once, like a failing find() of the reference
		rc = -1;
	}
	return rc;
}

// One lane: its own engine context on device d_device + lane % d_ndev; takes batches in submission order, scans
// and formats them concurrently with the other lanes, prints when it is the batch's turn.
void FileGrep::lane_main(int lane)
{
	Pipeline &p = *d_pipe;
	const int device = d_device + lane % d_ndev;
	gscan_ctx *ctx = nullptr;
	std::vector<gscan_unit> units;
	std::vector<gscan_match_view> view;
	std::vector<std::pair<uint32_t, std::string>> outs; // (file_seq, text of one window)
	const bool trace = g_trace;
	for (;;) {
		Batch *b = nullptr;
		{
			std::unique_lock<std::mutex> lk(p.mu);
			p.cv_job.wait(lk, [&] { return p.stop || !p.jobs.empty(); });
			if (p.jobs.empty()) break;
			b = p.jobs.front();
			p.jobs.pop_front();
		}
		p.cv_space.notify_one();

		std::string err;
		int rc = 0;
		if (!ctx) {
			if (!(ctx = gscan_open(device))) {
				err = std::string("FileGrep::find::gscan_open: ") + gscan_last_error();
				rc = -1;
			}
			trace_ts("engine context open", lane);
		}
		outs.clear();
		if (rc == 0) {
			units.resize(b->windows.size());
			for (size_t i = 0; i < units.size(); i++) {
				const bool by_fd = b->windows[i].fd >= 0;
				units[i].ptr = by_fd ? reinterpret_cast<const uint8_t *>((intptr_t)b->windows[i].fd) : b->windows[i].map;
				units[i].len = b->windows[i].clen;
				units[i].base_off = b->windows[i].off;
				units[i].file_id = (uint32_t)i; // index of the window in this batch
				units[i].flags = by_fd ? GSCAN_UNIT_FD : 0u;
			}
			uint32_t mode = GSCAN_MODE_ALL;
			if (d_print_line) mode = GSCAN_MODE_LINE;                 // resume after the printed line (grab.cc:188-209)
			else if (!d_print_offset) mode = GSCAN_MODE_FIRST;        // "matches" once per window (grab.cc:204-207)
			if (d_single_match) mode = GSCAN_MODE_FIRST;              // grab.cc:211-212
			gscan_match *matches = nullptr;
			size_t n = 0;
			rc = gscan_scan_batch(ctx, d_pat, units.data(), units.size(), mode, &matches, &n);
			if (rc < 0) err = std::string("FileGrep::find::scan: ") + gscan_why(ctx);
			else {
				gscan_stats st;
				static std::atomic<bool> warned{false};
				if (gscan_last_stats(ctx, &st) == 0 && st.vm_limit_hit && !warned.exchange(true))
					std::cerr << "grab-b200: backtracking limit reached in at least one window; the search of that window stopped there "
					             "(the reference's loop does the same when pcre_exec reports an error)\n";
			}
			if (trace) {
				gscan_stats st;
				gscan_last_stats(ctx, &st);
				char line[256];
				size_t nfd = 0;
				for (const Window &w : b->windows) nfd += w.fd >= 0 ? 1 : 0;
				snprintf(line, sizeof line, "gpu %d batch %llu: %zu windows (%zu as descriptors), %.1f MiB, staging+h2d %.2f ms, scan kernel %.3f ms, resolve %.3f ms, call %.2f ms, %zu matches",
				         device, (unsigned long long)b->seq, units.size(), nfd, (double)st.bytes_scanned / 1048576.0, st.h2d_ms,
				         st.scan_kernel_ms, st.resolve_ms, st.total_ms, n);
				trace_ts(line, lane);
			}
			// per window, in queue order: the text the reference would have flushed for it (grab.cc:175-213)
			size_t k = 0;
			for (size_t i = 0; i < b->windows.size() && rc == 0; i++) {
				view.clear();
				while (k < n && matches[k].file_id == (uint32_t)i) {
					view.push_back(gscan_match_view{matches[k].start, matches[k].match_len});
					k++;
				}
				if (view.empty()) continue;
				outs.emplace_back(b->windows[i].file_seq, std::string());
				format_window(b->windows[i], view.data(), view.size(), outs.back().second);
			}
			if (matches) gscan_free_matches(ctx, matches);
		}
		for (auto &w : b->windows) release(w); // grab.cc:215
		if (trace) trace_ts("windows unmapped", lane);

		{
			std::unique_lock<std::mutex> lk(p.mu);
			p.cv_turn.wait(lk, [&] { return p.next_out == b->seq; });
			// flush window by window under the stdout lock (grab.cc:217-234)
			for (auto &o : outs) {
				if (d_single_match && p.have_done && p.done_seq == o.first) continue; // grab.cc:232-233
				if (o.second.empty()) continue;
				pthread_mutex_lock(&g_stdout_lock);
				std::cout << o.second;
				pthread_mutex_unlock(&g_stdout_lock);
				if (d_single_match) { p.have_done = true; p.done_seq = o.first; }
			}
			if (!outs.empty()) {
				pthread_mutex_lock(&g_stdout_lock);
				std::cout.flush();
				pthread_mutex_unlock(&g_stdout_lock);
			}
			if (rc < 0 && !p.failed) { p.failed = true; p.err = err; }
			p.next_out++;
			p.inflight--;
		}
		p.cv_turn.notify_all();
		p.cv_idle.notify_all();
		delete b;
	}
	if (ctx) gscan_close(ctx);
	trace_ts("engine context closed", lane);
}

int FileGrep::find(const std::string &path)
{
	struct stat st;
	if (stat(path.c_str(), &st) < 0) {
		d_err = "FileGrep::find::stat: " + std::string(strerror(errno));
		return -1;
	}
	int r = 0;
	if (S_ISREG(st.st_mode)) r = find(path.c_str(), &st, FTW_F);
	else if (S_ISDIR(st.st_mode)) std::cerr << "Clever boy! Want recursion? Add -R!\n"; // grab.cc:253-254
	return r;
}

// nftw has no user pointer: like the reference (grab.cc:260-272) the walker reaches the object
// through a file-scope pointer, set for the duration of find_recursive()
static thread_local FileGrep *t_walker = nullptr;

static int walk_cb(const char *path, const struct stat *st, int typeflag, struct FTW *)
{
	if (typeflag == FTW_F && S_ISREG(st->st_mode)) {
		if (t_walker->find(path, st, typeflag) < 0) std::cerr << path << ": " << t_walker->why() << std::endl; // grab.cc:267-268
	}
	return 0;
}

int FileGrep::find_recursive(const std::string &path)
{
	d_recursive = true;
	t_walker = this;
	int r = nftw(path.c_str(), walk_cb, 1024, FTW_PHYS); // grab.cc:278
	t_walker = nullptr;
	if (flush() < 0) return -1;
	return r;
}

} // namespace grab_b200
