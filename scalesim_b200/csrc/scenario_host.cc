// scenario_host.cc — input at scale (no device work of its own): the TrafficSim file readers as a multi-threaded
// parser over mapped files, and the scenario image (one page-aligned binary file an engine starts from).
//
// The reference reads its scenario as text, line by line, on every rank: std::getline, boost::split into a vector of
// std::string, atol per field, then one boost::shared_ptr<Event/State> per line
// (src/trafficsim/traffic_sim.hpp:345-451: graph_read, road_read, trip_read). Field positions and the atol reading of
// numbers are kept exactly (see include/scalesim_b200.h); the mechanics are replaced:
//   * the file is mmap'ed and cut into `threads` pieces at line ends; every thread scans its piece twice with pointer
//     arithmetic (no string, no token vector): once to count what the piece yields, once to parse straight into the
//     final arrays at the piece's offsets — pieces keep file order, so line index = LP id still holds for the
//     partition and road files;
//   * the result is flat arrays (CSR for roads per LP and nodes per trip), which is what the device wants.
// The image is the packed form of those arrays: see ssb_scenario in the header.
// Everything here goes through the public C ABI only (ssb_scenario_load calls ssb_set_partition / ssb_set_model_data /
// ssb_load_states / ssb_load_events_raw), so the same file serves the device library and the developer emulation.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/scalesim_b200.h"
#include "ssb_host_internal.h"

namespace {

int fail(int code, const std::string& what) {
  ssb_internal_set_error(what);
  return code;
}

// read-only mapping of a whole file
struct mapped_file {
  const char* data = nullptr;
  size_t size = 0;
  bool ok = false;
  explicit mapped_file(const char* path) {
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return;
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); return; }
    size = (size_t)st.st_size;
    if (size == 0) { ok = true; ::close(fd); return; }
    void* p = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) return;
    ::madvise(p, size, MADV_SEQUENTIAL);
    data = (const char*)p;
    ok = true;
  }
  ~mapped_file() { if (data) ::munmap((void*)data, size); }
  mapped_file(const mapped_file&) = delete;
  void operator=(const mapped_file&) = delete;
};

// C atol over the token [b, e): blanks, optional sign, digits; anything else ends the number; no digits = 0
inline int64_t atol_token(const char* b, const char* e) {
  while (b < e && (*b == ' ' || (*b >= '\t' && *b <= '\r'))) ++b;
  bool neg = false;
  if (b < e && (*b == '-' || *b == '+')) { neg = *b == '-'; ++b; }
  uint64_t v = 0;
  while (b < e && *b >= '0' && *b <= '9') v = v * 10 + (uint64_t)(*b++ - '0');
  return neg ? -(int64_t)v : (int64_t)v;
}

// [begin, end) of piece t of nt: pieces start right after a '\n' (or at 0), so every line belongs to one piece
void piece(const mapped_file& f, int t, int nt, size_t* begin, size_t* end) {
  auto cut = [&](int k) -> size_t {
    if (k <= 0) return 0;
    if (k >= nt) return f.size;
    size_t p = f.size / (size_t)nt * (size_t)k;
    const void* nl = p < f.size ? std::memchr(f.data + p, '\n', f.size - p) : nullptr;
    return nl ? (size_t)((const char*)nl - f.data) + 1 : f.size;
  };
  *begin = cut(t);
  *end = cut(t + 1);
}

int thread_count(int asked, size_t bytes) {
  if (asked < 0) return std::min(-asked, 256);        // exactly that many pieces whatever the size (tests of the cutting)
  int n = asked > 0 ? asked : (int)std::thread::hardware_concurrency();
  if (n < 1) n = 1;
  const size_t by_size = bytes / (1u << 20) + 1;      // at least ~1 MiB of text per thread
  return (int)std::min<size_t>((size_t)std::min(n, 256), by_size);
}

// runs fn(t) for t in [0, nt) on nt threads
template <class F> void parallel(int nt, F fn) {
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(fn, t);
  fn(0);
  for (auto& x : th) x.join();
}

// calls line(b, e) for every line of [begin, end) (e excludes the '\n'; a last line without '\n' counts, as getline)
template <class F> void for_lines(const char* data, size_t begin, size_t end, F line) {
  const char* p = data + begin;
  const char* stop = data + end;
  while (p < stop) {
    const char* nl = (const char*)std::memchr(p, '\n', (size_t)(stop - p));
    const char* e = nl ? nl : stop;
    line(p, e);
    p = e + 1;
  }
}

int64_t* alloc64(int64_t n) { return (int64_t*)std::malloc((size_t)std::max<int64_t>(1, n) * sizeof(int64_t)); }

// One number of a ','-separated line starting at p (p < le): C atol reading (blanks, sign, digits; the rest of the
// field is ignored). Returns the position behind the field's ',' — or le + 1 when the field was the line's last.
inline const char* field(const char* p, const char* le, int64_t* out) {
  while (p < le && (*p == ' ' || (*p >= '\t' && *p <= '\r'))) ++p;
  bool neg = false;
  if (p < le && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
  uint64_t v = 0;
  while (p < le && (unsigned)(*p - '0') <= 9u) v = v * 10 + (uint64_t)(*p++ - '0');
  *out = neg ? -(int64_t)v : (int64_t)v;
  while (p < le && *p != ',') ++p;
  return p + 1;
}
// skips one field
inline const char* skip_field(const char* p, const char* le) {
  const char* c = (const char*)std::memchr(p, ',', (size_t)(le - p));
  return c ? c + 1 : le + 1;
}

// Two passes over every piece, both on all threads: the first counts what a piece yields (lines, roads / route nodes),
// the second parses straight into the final arrays at the piece's offsets — no per-thread growth, no concatenation.
struct piece_count { int64_t lines = 0, items = 0; };

// ---- traffic_reader::graph_read (traffic_sim.hpp:345-362) ----
int read_partition(const char* path, int threads, ssb_traffic_tables* out) {
  mapped_file f(path);
  if (!f.ok) return fail(SSB_ERR_INVALID, std::string("cannot open partition file ") + path + ": " + std::strerror(errno));
  const int nt = thread_count(threads, f.size);
  std::vector<piece_count> cnt((size_t)nt);
  parallel(nt, [&](int t) {
    size_t b, e;
    piece(f, t, nt, &b, &e);
    int64_t n = 0;
    for_lines(f.data, b, e, [&](const char*, const char*) { ++n; });
    cnt[(size_t)t].lines = n;
  });
  std::vector<int64_t> at((size_t)nt + 1, 0);
  for (int t = 0; t < nt; ++t) at[(size_t)t + 1] = at[(size_t)t] + cnt[(size_t)t].lines;
  out->n_part = at[(size_t)nt];
  if (!(out->part = alloc64(out->n_part))) return fail(SSB_ERR_INVALID, "out of host memory");
  parallel(nt, [&](int t) {
    size_t b, e;
    piece(f, t, nt, &b, &e);
    int64_t* o = out->part + at[(size_t)t];
    for_lines(f.data, b, e, [&](const char* lb, const char* le) { *o++ = atol_token(lb, le); });
  });
  return SSB_OK;
}

// ---- traffic_reader::road_read (traffic_sim.hpp:364-408), every partition's lines ----
// One line = one LP; roads split on ';', fields on ','. store = false: only counts the usable roads.
template <bool store>
inline int64_t road_line(const char* lb, const char* le, int64_t* sid, int64_t* dst, int64_t* speed, int64_t* len) {
  int64_t state_id = -1, n = 0;
  const char* rb = lb;
  for (;;) {
    const char* re = (const char*)std::memchr(rb, ';', (size_t)(le - rb));
    if (!re) re = le;
    // a road needs fields 0..6, i.e. at least six ',' (the reference indexes road[6])
    const char* p = rb;
    p = skip_field(p, re);
    if (p <= re) p = skip_field(p, re);
    if (p <= re) {
      int64_t f2, f3, f4, f5;
      const char* q = field(p, re, &f2);
      if (q <= re && (q = field(q, re, &f3)) <= re && (q = field(q, re, &f4)) <= re && (q = field(q, re, &f5)) <= re) {
        // q is at field 6 (possibly empty): the road counts
        if (store) { state_id = f2; dst[n] = f3; speed[n] = f4; len[n] = f5 != 0 ? f5 : 1; }
        ++n;
      }
    }
    if (re == le) break;
    rb = re + 1;
  }
  if (store) *sid = state_id;       // the LAST usable road of the line decides (quirk Q10)
  return n;
}

int read_roads(const char* path, int threads, ssb_traffic_tables* out) {
  mapped_file f(path);
  if (!f.ok) return fail(SSB_ERR_INVALID, std::string("cannot open road file ") + path + ": " + std::strerror(errno));
  const int nt = thread_count(threads, f.size);
  std::vector<piece_count> cnt((size_t)nt);
  parallel(nt, [&](int t) {
    size_t b, e;
    piece(f, t, nt, &b, &e);
    piece_count c;
    for_lines(f.data, b, e, [&](const char* lb, const char* le) {
      ++c.lines;
      c.items += road_line<false>(lb, le, nullptr, nullptr, nullptr, nullptr);
    });
    cnt[(size_t)t] = c;
  });
  std::vector<int64_t> line_at((size_t)nt + 1, 0), road_at((size_t)nt + 1, 0);
  for (int t = 0; t < nt; ++t) {
    line_at[(size_t)t + 1] = line_at[(size_t)t] + cnt[(size_t)t].lines;
    road_at[(size_t)t + 1] = road_at[(size_t)t] + cnt[(size_t)t].items;
  }
  out->n_states = line_at[(size_t)nt];
  out->n_roads = road_at[(size_t)nt];
  out->state_ids = alloc64(out->n_states);
  out->road_start = alloc64(out->n_states + 1);
  out->road_dst = alloc64(out->n_roads);
  out->road_speed = alloc64(out->n_roads);
  out->road_len = alloc64(out->n_roads);
  if (!out->state_ids || !out->road_start || !out->road_dst || !out->road_speed || !out->road_len)
    return fail(SSB_ERR_INVALID, "out of host memory");
  parallel(nt, [&](int t) {
    size_t b, e;
    piece(f, t, nt, &b, &e);
    int64_t line = line_at[(size_t)t], road = road_at[(size_t)t];
    for_lines(f.data, b, e, [&](const char* lb, const char* le) {
      out->road_start[line] = road;
      road += road_line<true>(lb, le, out->state_ids + line, out->road_dst + road, out->road_speed + road, out->road_len + road);
      ++line;
    });
  });
  out->road_start[out->n_states] = out->n_roads;
  return SSB_OK;
}

// ---- traffic_reader::trip_read (traffic_sim.hpp:410-451), every rank's lines ----
// fields 0..3, then the route (fields 4..); a line with fewer than 5 fields is skipped (returns -1).
// store = false: only counts the route's nodes.
template <bool store>
inline int64_t trip_line(const char* lb, const char* le, int64_t* id, int64_t* dep, int64_t* nodes) {
  const char* p = skip_field(lb, le);                   // field 0
  if (p > le) return -1;
  int64_t vid = 0, t = 0;
  if (store) p = field(p, le, &vid); else p = skip_field(p, le);      // field 1: vehicle id
  if (p > le) return -1;
  p = skip_field(p, le);                                // field 2: ignored
  if (p > le) return -1;
  if (store) p = field(p, le, &t); else p = skip_field(p, le);        // field 3: arrival = departure
  if (p > le) return -1;
  int64_t n = 0;
  if (store) {
    *id = vid; *dep = t;
    while (p <= le) p = field(p, le, &nodes[n++]);
  } else {
    n = 1;
    for (const char* q = p; q < le; ++q) n += *q == ',';
  }
  return n;
}

int read_trips(const char* path, int threads, ssb_traffic_tables* out) {
  mapped_file f(path);
  if (!f.ok) return fail(SSB_ERR_INVALID, std::string("cannot open trip file ") + path + ": " + std::strerror(errno));
  const int nt = thread_count(threads, f.size);
  std::vector<piece_count> cnt((size_t)nt);
  parallel(nt, [&](int t) {
    size_t b, e;
    piece(f, t, nt, &b, &e);
    piece_count c;
    for_lines(f.data, b, e, [&](const char* lb, const char* le) {
      const int64_t n = trip_line<false>(lb, le, nullptr, nullptr, nullptr);
      if (n >= 0) { ++c.lines; c.items += n; }
    });
    cnt[(size_t)t] = c;
  });
  std::vector<int64_t> trip_at((size_t)nt + 1, 0), node_at((size_t)nt + 1, 0);
  for (int t = 0; t < nt; ++t) {
    trip_at[(size_t)t + 1] = trip_at[(size_t)t] + cnt[(size_t)t].lines;
    node_at[(size_t)t + 1] = node_at[(size_t)t] + cnt[(size_t)t].items;
  }
  out->n_trips = trip_at[(size_t)nt];
  out->n_nodes = node_at[(size_t)nt];
  out->trip_id = alloc64(out->n_trips);
  out->trip_dep = alloc64(out->n_trips);
  out->route_start = alloc64(out->n_trips + 1);
  out->route_nodes = alloc64(out->n_nodes);
  if (!out->trip_id || !out->trip_dep || !out->route_start || !out->route_nodes) return fail(SSB_ERR_INVALID, "out of host memory");
  parallel(nt, [&](int t) {
    size_t b, e;
    piece(f, t, nt, &b, &e);
    int64_t trip = trip_at[(size_t)t], node = node_at[(size_t)t];
    for_lines(f.data, b, e, [&](const char* lb, const char* le) {
      const int64_t n = trip_line<true>(lb, le, out->trip_id + trip, out->trip_dep + trip, out->route_nodes + node);
      if (n < 0) return;
      out->route_start[trip++] = node;
      node += n;
    });
  });
  out->route_start[out->n_trips] = out->n_nodes;
  return SSB_OK;
}

// ---- scenario image file ----
const char kMagic[8] = {'S', 'S', 'B', 'S', 'C', 'N', '0', '1'};
const size_t kAlign = 4096;
// header: magic, then 8-byte little-endian integers
struct image_header {
  char magic[8];
  int64_t model, state_words, n_lp, finish_time, n_part, n_states, n_events;
  int64_t slot_bytes[SSB_SCENARIO_SLOTS];
  int64_t off_part, off_state_lps, off_states, off_slot[SSB_SCENARIO_SLOTS], off_events, file_bytes;
};
static_assert(sizeof(image_header) <= kAlign, "header fits the first page");

size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

bool write_at(int fd, size_t off, const void* p, size_t n) {
  const char* c = (const char*)p;
  while (n) {
    ssize_t w = ::pwrite(fd, c, std::min<size_t>(n, (size_t)1 << 30), (off_t)off);
    if (w <= 0) return false;
    c += w; off += (size_t)w; n -= (size_t)w;
  }
  return true;
}

inline uint32_t owner_of(const ssb_scenario* sc, int64_t lp, int world) {
  if (sc->n_part == 0) return (uint32_t)(lp % world);
  return lp < sc->n_part ? (uint32_t)(sc->lp_to_part[lp] % world) : 0u;
}

}  // namespace

extern "C" {

int ssb_traffic_read(const char* partition_file, const char* road_file, const char* trip_file, int threads,
                     ssb_traffic_tables* out) {
  if (!out) return fail(SSB_ERR_INVALID, "null argument");
  std::memset(out, 0, sizeof(*out));
  int rc = SSB_OK;
  if (partition_file && (rc = read_partition(partition_file, threads, out))) { ssb_traffic_free(out); return rc; }
  if (road_file && (rc = read_roads(road_file, threads, out))) { ssb_traffic_free(out); return rc; }
  if (trip_file && (rc = read_trips(trip_file, threads, out))) { ssb_traffic_free(out); return rc; }
  return SSB_OK;
}

void ssb_traffic_free(ssb_traffic_tables* t) {
  if (!t) return;
  int64_t** all[] = {&t->part, &t->state_ids, &t->road_start, &t->road_dst, &t->road_speed, &t->road_len,
                     &t->trip_id, &t->trip_dep, &t->route_start, &t->route_nodes};
  for (int64_t** p : all) { std::free(*p); *p = nullptr; }
  t->n_part = t->n_states = t->n_roads = t->n_trips = t->n_nodes = 0;
}

int ssb_traffic_write_csv(const ssb_traffic_tables* t, const char* partition_file, const char* road_file, const char* trip_file) {
  if (!t) return fail(SSB_ERR_INVALID, "null argument");
  auto open = [&](const char* path) -> std::FILE* {
    std::FILE* f = std::fopen(path, "wb");
    if (f) std::setvbuf(f, nullptr, _IOFBF, 1 << 22);
    return f;
  };
  if (partition_file) {
    std::FILE* f = open(partition_file);
    if (!f) return fail(SSB_ERR_INVALID, std::string("cannot write ") + partition_file);
    for (int64_t i = 0; i < t->n_part; ++i) std::fprintf(f, "%lld\n", (long long)t->part[i]);
    if (std::fclose(f) != 0) return fail(SSB_ERR_INVALID, std::string("cannot write ") + partition_file);
  }
  if (road_file) {
    std::FILE* f = open(road_file);
    if (!f) return fail(SSB_ERR_INVALID, std::string("cannot write ") + road_file);
    for (int64_t i = 0; i < t->n_states; ++i) {
      for (int64_t r = t->road_start[i]; r < t->road_start[i + 1]; ++r)
        std::fprintf(f, "%sR,%lld,%lld,%lld,%lld,%lld,1", r == t->road_start[i] ? "" : ";", (long long)r, (long long)t->state_ids[i],
                     (long long)t->road_dst[r], (long long)t->road_speed[r], (long long)t->road_len[r]);
      std::fputc('\n', f);
    }
    if (std::fclose(f) != 0) return fail(SSB_ERR_INVALID, std::string("cannot write ") + road_file);
  }
  if (trip_file) {
    std::FILE* f = open(trip_file);
    if (!f) return fail(SSB_ERR_INVALID, std::string("cannot write ") + trip_file);
    for (int64_t i = 0; i < t->n_trips; ++i) {
      std::fprintf(f, "%lld,%lld,0,%lld", (long long)i, (long long)t->trip_id[i], (long long)t->trip_dep[i]);
      for (int64_t k = t->route_start[i]; k < t->route_start[i + 1]; ++k) std::fprintf(f, ",%lld", (long long)t->route_nodes[k]);
      std::fputc('\n', f);
    }
    if (std::fclose(f) != 0) return fail(SSB_ERR_INVALID, std::string("cannot write ") + trip_file);
  }
  return SSB_OK;
}

int ssb_scenario_save(const char* path, const ssb_scenario* sc) {
  if (!path || !sc) return fail(SSB_ERR_INVALID, "null argument");
  if (sc->n_part < 0 || sc->n_states < 0 || sc->n_events < 0 || sc->state_words < 0 || sc->n_lp < 0 ||
      (sc->n_part && !sc->lp_to_part) || (sc->n_states && (!sc->state_lps || !sc->states)) || (sc->n_events && !sc->events))
    return fail(SSB_ERR_INVALID, "inconsistent scenario description");
  image_header h;
  std::memset(&h, 0, sizeof(h));
  std::memcpy(h.magic, kMagic, 8);
  h.model = sc->model; h.state_words = sc->state_words; h.n_lp = sc->n_lp; h.finish_time = sc->finish_time;
  h.n_part = sc->n_part; h.n_states = sc->n_states; h.n_events = sc->n_events;
  size_t off = kAlign;
  auto place = [&](size_t bytes) { size_t at = off; off = align_up(off + bytes); return (int64_t)at; };
  h.off_part = place((size_t)sc->n_part * 8);
  h.off_state_lps = place((size_t)sc->n_states * 8);
  h.off_states = place((size_t)sc->n_states * (size_t)sc->state_words * 4);
  for (int s = 0; s < SSB_SCENARIO_SLOTS; ++s) {
    if (sc->slot_bytes[s] < 0 || (sc->slot_bytes[s] && !sc->slot_data[s])) return fail(SSB_ERR_INVALID, "inconsistent model data slot");
    h.slot_bytes[s] = sc->slot_bytes[s];
    h.off_slot[s] = place((size_t)sc->slot_bytes[s]);
  }
  h.off_events = place((size_t)sc->n_events * sizeof(ssb_raw_record));
  h.file_bytes = (int64_t)off;
  const std::string tmp = std::string(path) + ".tmp";
  int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return fail(SSB_ERR_INVALID, "cannot write " + tmp + ": " + std::strerror(errno));
  bool ok = ::ftruncate(fd, (off_t)off) == 0 && write_at(fd, 0, &h, sizeof(h)) &&
            write_at(fd, (size_t)h.off_part, sc->lp_to_part, (size_t)sc->n_part * 8) &&
            write_at(fd, (size_t)h.off_state_lps, sc->state_lps, (size_t)sc->n_states * 8) &&
            write_at(fd, (size_t)h.off_states, sc->states, (size_t)sc->n_states * (size_t)sc->state_words * 4) &&
            write_at(fd, (size_t)h.off_events, sc->events, (size_t)sc->n_events * sizeof(ssb_raw_record));
  for (int s = 0; ok && s < SSB_SCENARIO_SLOTS; ++s) ok = write_at(fd, (size_t)h.off_slot[s], sc->slot_data[s], (size_t)sc->slot_bytes[s]);
  ok = (::close(fd) == 0) && ok;
  if (!ok || std::rename(tmp.c_str(), path) != 0) {
    ::unlink(tmp.c_str());
    return fail(SSB_ERR_INVALID, std::string("cannot write ") + path + ": " + std::strerror(errno));
  }
  return SSB_OK;
}

int ssb_scenario_open(const char* path, ssb_scenario* sc) {
  if (!path || !sc) return fail(SSB_ERR_INVALID, "null argument");
  std::memset(sc, 0, sizeof(*sc));
  int fd = ::open(path, O_RDONLY);
  if (fd < 0) return fail(SSB_ERR_INVALID, std::string("cannot open scenario image ") + path + ": " + std::strerror(errno));
  struct stat st;
  if (::fstat(fd, &st) != 0 || (size_t)st.st_size < kAlign) { ::close(fd); return fail(SSB_ERR_INVALID, std::string(path) + ": not a scenario image (too short)"); }
  const size_t size = (size_t)st.st_size;
  void* p = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  ::close(fd);
  if (p == MAP_FAILED) return fail(SSB_ERR_INVALID, std::string("mmap ") + path + ": " + std::strerror(errno));
  const char* base = (const char*)p;
  image_header h;
  std::memcpy(&h, base, sizeof(h));
  auto bad = [&](const char* why) { ::munmap(p, size); return fail(SSB_ERR_INVALID, std::string(path) + ": " + why); };
  if (std::memcmp(h.magic, kMagic, 8) != 0) return bad("not a scenario image (magic)");
  if (h.file_bytes != (int64_t)size) return bad("truncated or grown (size in the header differs)");
  if (h.n_part < 0 || h.n_states < 0 || h.n_events < 0 || h.state_words < 0 || h.state_words > 64 || h.n_lp < 0) return bad("negative count");
  // every section inside the file (counts are bounded by the size first, so the products cannot overflow)
  auto inside = [&](int64_t off, int64_t count, int64_t elem) {
    return off >= (int64_t)kAlign && off % 8 == 0 && count <= (int64_t)size && (uint64_t)off + (uint64_t)count * (uint64_t)elem <= size;
  };
  if (!inside(h.off_part, h.n_part, 8) || !inside(h.off_state_lps, h.n_states, 8) ||
      !inside(h.off_states, h.n_states, h.state_words * 4) || !inside(h.off_events, h.n_events, (int64_t)sizeof(ssb_raw_record)))
    return bad("section outside the file");
  for (int s = 0; s < SSB_SCENARIO_SLOTS; ++s)
    if (h.slot_bytes[s] < 0 || !inside(h.off_slot[s], h.slot_bytes[s], 1)) return bad("model data slot outside the file");
  sc->model = (int32_t)h.model; sc->state_words = (int32_t)h.state_words; sc->n_lp = h.n_lp; sc->finish_time = h.finish_time;
  sc->n_part = h.n_part; sc->lp_to_part = (const int64_t*)(base + h.off_part);
  sc->n_states = h.n_states; sc->state_lps = (const int64_t*)(base + h.off_state_lps); sc->states = (const uint32_t*)(base + h.off_states);
  for (int s = 0; s < SSB_SCENARIO_SLOTS; ++s) { sc->slot_bytes[s] = h.slot_bytes[s]; sc->slot_data[s] = h.slot_bytes[s] ? base + h.off_slot[s] : nullptr; }
  sc->n_events = h.n_events; sc->events = (const ssb_raw_record*)(base + h.off_events);
  sc->mapping = p; sc->mapping_bytes = (int64_t)size;
  return SSB_OK;
}

void ssb_scenario_close(ssb_scenario* sc) {
  if (!sc) return;
  if (sc->mapping) ::munmap(sc->mapping, (size_t)sc->mapping_bytes);
  std::memset(sc, 0, sizeof(*sc));
}

int ssb_scenario_load(ssb_engine* e, const ssb_scenario* sc) {
  if (!e || !sc) return SSB_ERR_INVALID;
  ssb_config cfg;
  int rc = ssb_get_config(e, &cfg);
  if (rc) return rc;
  auto efail = [&](int code, const std::string& what) { ssb_internal_set_engine_error(e, what); return code; };
  if (cfg.model != sc->model) return efail(SSB_ERR_MODEL, "scenario image is for model " + std::to_string(sc->model) + ", engine runs model " + std::to_string(cfg.model));
  if (cfg.n_lp != sc->n_lp) return efail(SSB_ERR_INVALID, "scenario image has " + std::to_string(sc->n_lp) + " LPs, engine " + std::to_string(cfg.n_lp));
  if (sc->state_words != ssb_state_words(e)) return efail(SSB_ERR_MODEL, "scenario image has another state layout");
  if (sc->n_part > sc->n_lp) return efail(SSB_ERR_INVALID, "partition table longer than n_lp");
  const int world = cfg.world, rank = cfg.rank;
  // partition numbers reduced modulo world (runner.hpp:98 + traffic_sim.hpp:377); LPs past the table's end go to 0
  if (sc->n_part) {
    std::vector<int64_t> part((size_t)sc->n_lp, 0);
    for (int64_t i = 0; i < sc->n_part; ++i) {
      if (sc->lp_to_part[i] < 0) return efail(SSB_ERR_INVALID, "negative partition number");
      part[(size_t)i] = sc->lp_to_part[i] % world;
    }
    if ((rc = ssb_set_partition(e, part.data(), sc->n_lp))) return rc;
  } else if ((rc = ssb_set_partition(e, nullptr, sc->n_lp))) {
    return rc;
  }
  for (int s = 0; s < SSB_SCENARIO_SLOTS; ++s)
    if (sc->slot_bytes[s] && (rc = ssb_set_model_data(e, s, sc->slot_data[s], (size_t)sc->slot_bytes[s]))) return rc;
  const size_t sw = (size_t)sc->state_words;
  if (world == 1) {
    if ((rc = ssb_load_states(e, sc->state_lps, sc->states, sc->n_states))) return rc;
    return ssb_load_events_raw(e, sc->events, sc->n_events);
  }
  {
    std::vector<int64_t> ids;
    std::vector<uint32_t> words;
    for (int64_t i = 0; i < sc->n_states; ++i) {
      const int64_t lp = sc->state_lps[i];
      if (lp < 0 || lp >= sc->n_lp) return efail(SSB_ERR_RANGE, "state of an LP outside [0, n_lp)");
      if (owner_of(sc, lp, world) != (uint32_t)rank) continue;
      ids.push_back(lp);
      words.insert(words.end(), sc->states + (size_t)i * sw, sc->states + (size_t)(i + 1) * sw);
    }
    if ((rc = ssb_load_states(e, ids.data(), words.data(), (int64_t)ids.size()))) return rc;
  }
  // the events of this rank's LPs, filtered through a bounded buffer
  const int64_t kChunk = 1 << 20;
  std::vector<ssb_raw_record> mine;
  mine.reserve((size_t)kChunk);
  for (int64_t i = 0; i < sc->n_events; ++i) {
    const ssb_raw_record& r = sc->events[i];
    if ((int64_t)r.dst >= sc->n_lp) return efail(SSB_ERR_RANGE, "event for an LP outside [0, n_lp)");
    if (owner_of(sc, (int64_t)r.dst, world) != (uint32_t)rank) continue;
    mine.push_back(r);
    if ((int64_t)mine.size() == kChunk) {
      if ((rc = ssb_load_events_raw(e, mine.data(), kChunk))) return rc;
      mine.clear();
    }
  }
  return ssb_load_events_raw(e, mine.data(), (int64_t)mine.size());
}

int ssb_traffic_compile(const char* partition_file, const char* road_file, const char* trip_file, int64_t finish_time,
                        int threads, const char* image_path) {
  if (!road_file || !trip_file || !image_path) return fail(SSB_ERR_INVALID, "null argument");
  ssb_traffic_tables t;
  int rc = ssb_traffic_read(partition_file, road_file, trip_file, threads, &t);
  if (rc) return rc;
  struct guard { ssb_traffic_tables* t; ~guard() { ssb_traffic_free(t); } } g = {&t};
  const int kInline = 2, kWords = 2 + 3 * kInline;      // TrafficModel::kInlineRoads (csrc/ssb_models.cuh)
  auto fits = [](int64_t v) { return v >= 0 && v < 0xFFFFFFFFll; };
  // n_lp as runner::init sees it: the partition table's length; without one, the largest id mentioned anywhere + 1
  const int nt = thread_count(threads, (size_t)(t.n_nodes + t.n_roads) * 8);
  auto span = [&](int64_t n, int k, int64_t* a, int64_t* b) { *a = n * k / nt; *b = n * (k + 1) / nt; };
  std::vector<int64_t> top((size_t)nt, 0);
  parallel(nt, [&](int k) {
    int64_t m = 0, a, b;
    span(t.n_states, k, &a, &b);
    for (int64_t i = a; i < b; ++i) m = std::max(m, t.state_ids[i] + 1);
    span(t.n_roads, k, &a, &b);
    for (int64_t i = a; i < b; ++i) m = std::max(m, t.road_dst[i] + 1);
    span(t.n_nodes, k, &a, &b);
    for (int64_t i = a; i < b; ++i) m = std::max(m, t.route_nodes[i] + 1);
    top[(size_t)k] = m;
  });
  int64_t n_lp = t.n_part;
  for (int64_t m : top) n_lp = std::max(n_lp, m);
  // LP states (scalesim/models/traffic_model.hpp pack_state): [id, n, up to two roads inline | index into the road table]
  std::vector<uint32_t> words((size_t)t.n_states * kWords, 0u);
  std::vector<int64_t> road_table;
  for (int64_t i = 0; i < t.n_states; ++i) {
    uint32_t* w = &words[(size_t)i * kWords];
    const int64_t a = t.road_start[i], n = t.road_start[i + 1] - a;
    // a line without a usable road yields state id -1: such an LP cannot be addressed (the reference dies on it)
    if (t.state_ids[i] < 0 || !fits(t.state_ids[i])) return fail(SSB_ERR_RANGE, "road file line " + std::to_string(i + 1) + ": no usable road / LP id out of range");
    w[0] = (uint32_t)t.state_ids[i]; w[1] = (uint32_t)n;
    for (int64_t k = a; k < a + n; ++k)
      if (!fits(t.road_dst[k]) || !fits(t.road_speed[k]) || !fits(t.road_len[k])) return fail(SSB_ERR_RANGE, "road file line " + std::to_string(i + 1) + ": value does not fit 32 bits");
    if (n <= kInline) {
      for (int64_t k = 0; k < n; ++k) { w[2 + 3 * k] = (uint32_t)t.road_dst[a + k]; w[3 + 3 * k] = (uint32_t)t.road_speed[a + k]; w[4 + 3 * k] = (uint32_t)t.road_len[a + k]; }
    } else {
      w[2] = (uint32_t)(road_table.size() / 3);
      for (int64_t k = a; k < a + n; ++k) { road_table.push_back(t.road_dst[k]); road_table.push_back(t.road_speed[k]); road_table.push_back(t.road_len[k]); }
    }
  }
  // one initial event per trip (pack_event): source -1, destination = first node, p0 = offset in the pool, p1 = nodes
  if (t.n_nodes >= 0xFFFFFFFFll) return fail(SSB_ERR_RANGE, "route pool too large");
  ssb_raw_record* ev = (ssb_raw_record*)std::malloc(std::max<size_t>(1, (size_t)t.n_trips) * sizeof(ssb_raw_record));
  if (!ev) return fail(SSB_ERR_INVALID, "out of host memory");
  struct ev_guard { ssb_raw_record* p; ~ev_guard() { std::free(p); } } eg = {ev};
  std::vector<int64_t> bad_line((size_t)nt, -1);
  parallel(nt, [&](int k) {
    int64_t lo, hi;
    span(t.n_trips, k, &lo, &hi);
    for (int64_t i = lo; i < hi; ++i) {
      const int64_t a = t.route_start[i], n = t.route_start[i + 1] - a;
      if (!fits(t.trip_id[i]) || t.trip_dep[i] < 0 || t.trip_dep[i] >= (1ll << 31) || !fits(t.route_nodes[a])) {
        if (bad_line[(size_t)k] < 0) bad_line[(size_t)k] = i + 1;
        continue;
      }
      ssb_raw_record& r = ev[i];
      r.key = ((uint64_t)t.trip_dep[i] << 32) | (uint64_t)t.trip_id[i];
      r.dst = (uint32_t)t.route_nodes[a]; r.src = 0xFFFFFFFFu; r.send_time = (uint32_t)t.trip_dep[i];
      r.p0 = (uint32_t)a; r.p1 = (uint32_t)n; r.tag = 0;
    }
  });
  for (int64_t b : bad_line)
    if (b >= 0) return fail(SSB_ERR_RANGE, "trip file line " + std::to_string(b) + ": value out of the device record's range");
  const int64_t zero = 0;
  ssb_scenario sc;
  std::memset(&sc, 0, sizeof(sc));
  sc.model = SSB_MODEL_TRAFFIC; sc.state_words = kWords; sc.n_lp = n_lp; sc.finish_time = finish_time;
  sc.n_part = t.n_part; sc.lp_to_part = t.part;
  sc.n_states = t.n_states; sc.state_lps = t.state_ids; sc.states = words.data();
  sc.slot_data[0] = t.n_nodes ? t.route_nodes : &zero; sc.slot_bytes[0] = (t.n_nodes ? t.n_nodes : 1) * 8;
  if (!road_table.empty()) { sc.slot_data[1] = road_table.data(); sc.slot_bytes[1] = (int64_t)road_table.size() * 8; }
  sc.n_events = t.n_trips; sc.events = ev;
  return ssb_scenario_save(image_path, &sc);
}

}  // extern "C"
