// filegrep.h -- host-side mirror of the reference's operator interface for the scan path: class
// FileGrep of /root/reference/src/grab.h:39-85, same method names, argument meaning and error
// behaviour (int 0 / -1 + why()), with the per-chunk match loop (grab.cc:175-213) delegated to the
// CUDA engine through the C ABI (include/gscan.h).  Everything else the reference does on the CPU
// stays on the CPU here too: stat / open / mmap windows with 4 KiB overlap (grab.cc:131-169), the
// small-file skip (:133-135), output formatting (:182-207) and the locked per-chunk flush (:217-234).
//
// One deliberate difference in mechanics, none in output: windows are queued and handed to the GPU
// in batches (one kernel launch per ~256 MiB instead of one pcre_exec per match); matches come
// back sorted by (window, offset), so the bytes written to stdout are the reference's, in the
// reference's order.
//
// Batches are scanned by "lanes": worker threads that each own one engine context on one GPU
// (SURVEY.md 8(f) f1/f4).  The walking thread keeps stat/open/mmap-ing while the lanes stage, scan
// and format; a sequencer prints the batches in submission order, so stdout does not depend on the
// number of lanes or GPUs.  Since the batches of one huge file are the reference's own windows
// (grab.cc:151-159, 4 KiB overlap), spreading them over several GPUs reproduces Q3 exactly.
#pragma once

#include <sys/stat.h>

#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

struct gscan_ctx;
struct gscan_pattern;

namespace grab_b200 {

class FileGrep {
public:
	FileGrep();
	explicit FileGrep(int device);
	~FileGrep();

	const char *why() { return d_err.c_str(); }                    // grab.h:61-64
	void recurse() { d_recursive = true; }                         // grab.h:66-69
	void show_path(bool b) { d_print_path = b; }                   // grab.h:71-74
	int prepare(const std::string &regex);                         // grab.cc:101-123
	void config(const std::map<std::string, size_t> &config);      // grab.cc:83-98
	int find(const std::string &path);                             // grab.cc:242-257
	int find(const char *path, const struct stat *st, int typeflag); // grab.cc:131-239
	int find_recursive(const std::string &path);                   // grab.cc:275-279

	// engine-side additions (not in the reference): queued windows are scanned and printed here;
	// called automatically when the queue is full and from the destructor
	int flush();
	void literal(bool b) { d_literal = b; }        // -S
	void strict_reference(bool b) { d_strict = b; } // Q2 on/off (default on: bit-identical output)
	void batch_bytes(size_t n) { d_batch_bytes = n; }
	void devices(int first, int count) { d_device = first; d_ndev = count < 1 ? 1 : count; } // lanes go to first..first+count-1
	void lanes_per_device(int n) { d_lanes_per_dev = n < 1 ? 1 : n; }
	// descriptors this object may keep open at once (queued descriptor windows); 0: map the windows like the reference
	void fd_budget(size_t n) { d_fd_budget = n; }
	int minlen() const { return d_minlen; }

private:
	struct Window {
		std::string path;
		uint8_t *map = nullptr;   // mmap'd window (grab.cc:161); null for a descriptor window
		int fd = -1;              // descriptor window: read by the engine's staging threads (GSCAN_UNIT_FD), closed in release()
		size_t clen = 0;          // window length
		uint64_t off = 0;         // file offset of the window
		uint32_t file_seq = 0;    // which find() call the window belongs to (for -s early exit)
	};

	struct Batch;
	struct Pipeline;
	friend struct Pipeline;
	int submit();                      // hand the queued windows to the lanes (blocks only when they are all busy)
	void flush_quietly();              // submit + wait, leaving a lane failure for flush() to report
	size_t max_windows_per_batch() const;
	void lane_main(int lane);
	void format_window(const Window &w, const struct gscan_match_view *m, size_t n, std::string &out) const;
	static void release(Window &w);

	std::string d_err;
	int d_minlen = 1;
	bool d_print_line = true, d_print_offset = false, d_recursive = false, d_colored = false, d_print_path = false,
	     d_single_match = false, d_low_mem = false, d_literal = false, d_strict = true;
	size_t d_chunk_size = (size_t)1 << 30;   // grab.h:48
	uid_t d_my_uid = 0;
	int d_device = 0, d_ndev = 1, d_lanes_per_dev = 1;
	gscan_pattern *d_pat = nullptr;
	std::unique_ptr<Pipeline> d_pipe;
	std::vector<Window> d_queue;
	size_t d_queued_bytes = 0;
	size_t d_batch_bytes = (size_t)256 << 20;
	size_t d_fd_budget = 0;
	uint32_t d_file_seq = 0;
};

} // namespace grab_b200
