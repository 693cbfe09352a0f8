// host.cpp — FASTA staging and the synthetic protein-set generator (include/kc_host.h).
// Host code only; compiled with the engine into libkc_b200.so.
#include <cuda_runtime_api.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/kc_b200.h"
#include "../../include/kc_host.h"

struct kc_fasta {
  uint64_t n = 0, n_res = 0, n_missing = 0;
  uint8_t* residues = nullptr;
  bool pinned = false;
  std::vector<uint64_t> offsets;
  std::vector<uint32_t> class_ids;
  std::vector<std::string> class_names;
  std::vector<std::string> ids;
};

namespace {

template <class F>
void parallel_chunks(int threads, uint64_t n, uint64_t grain, F f) {
  if (threads <= 1 || n <= grain) {
    f(0, n);
    return;
  }
  std::atomic<uint64_t> cursor{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] {
      for (;;) {
        const uint64_t lo = cursor.fetch_add(grain);
        if (lo >= n) break;
        f(lo, std::min(n, lo + grain));
      }
    });
  for (auto& th : pool) th.join();
}

uint8_t* alloc_residues(uint64_t bytes, bool* pinned) {
  void* p = nullptr;
  *pinned = false;
  if (bytes == 0) bytes = 1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && cudaMallocHost(&p, bytes) == cudaSuccess) {
    *pinned = true;
    return static_cast<uint8_t*>(p);
  }
  cudaGetLastError();
  return static_cast<uint8_t*>(std::malloc(bytes));
}

int parse(const char* data, uint64_t len, int threads, kc_fasta** out) {
  if (threads < 1) threads = 1;
  kc_fasta* f = new kc_fasta();
  // record starts: '>' at the beginning of a line
  std::vector<uint64_t> starts;
  {
    const int T = threads;
    std::vector<std::vector<uint64_t>> local(T);
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t)
      pool.emplace_back([&, t] {
        const uint64_t lo = len * t / T, hi = len * (t + 1) / T;
        const char* p = data + lo;
        const char* end = data + hi;
        while (p < end) {
          const char* q = static_cast<const char*>(std::memchr(p, '>', end - p));
          if (!q) break;
          const uint64_t i = q - data;
          if (i == 0 || data[i - 1] == '\n') local[t].push_back(i);
          p = q + 1;
        }
      });
    for (auto& th : pool) th.join();
    for (auto& v : local) starts.insert(starts.end(), v.begin(), v.end());
  }
  const uint64_t n = starts.size();
  f->n = n;
  f->offsets.assign(n + 1, 0);
  f->ids.resize(n);
  std::vector<uint64_t> seq_begin(n), rec_end(n);
  parallel_chunks(threads, n, 1024, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t r = lo; r < hi; ++r) {
      const uint64_t s = starts[r], e = r + 1 < n ? starts[r + 1] : len;
      const char* nl = static_cast<const char*>(std::memchr(data + s, '\n', e - s));
      const uint64_t hdr_end = nl ? (uint64_t)(nl - data) : e;
      uint64_t id_end = s + 1;
      while (id_end < hdr_end && data[id_end] != ' ' && data[id_end] != '\t' && data[id_end] != '\r') ++id_end;
      f->ids[r].assign(data + s + 1, id_end - s - 1);
      seq_begin[r] = nl ? hdr_end + 1 : e;
      rec_end[r] = e;
      uint64_t cnt = 0;
      for (uint64_t i = seq_begin[r]; i < e; ++i) cnt += data[i] != '\n' && data[i] != '\r';
      f->offsets[r + 1] = cnt;
    }
  });
  for (uint64_t r = 0; r < n; ++r) f->offsets[r + 1] += f->offsets[r];
  f->n_res = f->offsets[n];
  f->residues = alloc_residues(f->n_res + 64, &f->pinned);
  if (!f->residues) {
    delete f;
    return KC_ENOMEM;
  }
  parallel_chunks(threads, n, 1024, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t r = lo; r < hi; ++r) {
      uint8_t* dst = f->residues + f->offsets[r];
      for (uint64_t i = seq_begin[r]; i < rec_end[r]; ++i) {
        const char c = data[i];
        if (c != '\n' && c != '\r') *dst++ = (uint8_t)c;
      }
    }
  });
  // class dictionary in first-occurrence order (Protein::get_amr_class, src/protein.rs:135-138)
  f->class_ids.resize(n);
  std::unordered_map<std::string, uint32_t> table;
  for (uint64_t r = 0; r < n; ++r) {
    const std::string& id = f->ids[r];
    std::string name;
    bool found = false;
    {
      // Rust split_terminator('|'): split on '|', drop one trailing empty piece
      std::vector<std::pair<size_t, size_t>> pieces;
      size_t pos = 0;
      for (;;) {
        const size_t bar = id.find('|', pos);
        if (bar == std::string::npos) {
          pieces.emplace_back(pos, id.size());
          break;
        }
        pieces.emplace_back(pos, bar);
        pos = bar + 1;
      }
      if (!pieces.empty() && pieces.back().first == pieces.back().second) pieces.pop_back();
      if (pieces.size() > 3) {
        name = id.substr(pieces[3].first, pieces[3].second - pieces[3].first);
        found = true;
      }
    }
    if (!found) f->n_missing++;
    auto it = table.find(name);
    if (it == table.end()) {
      it = table.emplace(name, (uint32_t)f->class_names.size()).first;
      f->class_names.push_back(name);
    }
    f->class_ids[r] = it->second;
  }
  *out = f;
  return KC_OK;
}

}  // namespace

extern "C" {

int kc_fasta_parse_buffer(const char* data, uint64_t len, int threads, kc_fasta** out) {
  if (!out || (len && !data)) return KC_EINVAL;
  return parse(data, len, threads, out);
}

int kc_fasta_parse_file(const char* path, int threads, kc_fasta** out) {
  if (!path || !out) return KC_EINVAL;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return KC_EINVAL;
  struct stat st;
  if (fstat(fd, &st) != 0) {
    close(fd);
    return KC_EINVAL;
  }
  const uint64_t len = (uint64_t)st.st_size;
  if (len == 0) {
    close(fd);
    return parse("", 0, threads, out);
  }
  void* m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return KC_ENOMEM;
  madvise(m, len, MADV_SEQUENTIAL);
  const int rc = parse(static_cast<const char*>(m), len, threads, out);
  munmap(m, len);
  return rc;
}

void kc_fasta_free(kc_fasta* f) {
  if (!f) return;
  if (f->residues) {
    if (f->pinned) cudaFreeHost(f->residues); else std::free(f->residues);
  }
  delete f;
}
uint64_t kc_fasta_n_proteins(const kc_fasta* f) { return f->n; }
uint64_t kc_fasta_n_residues(const kc_fasta* f) { return f->n_res; }
const uint8_t* kc_fasta_residues(const kc_fasta* f) { return f->residues; }
const uint64_t* kc_fasta_offsets(const kc_fasta* f) { return f->offsets.data(); }
const uint32_t* kc_fasta_class_ids(const kc_fasta* f) { return f->class_ids.data(); }
uint32_t kc_fasta_n_classes(const kc_fasta* f) { return (uint32_t)f->class_names.size(); }
uint64_t kc_fasta_n_missing_class(const kc_fasta* f) { return f->n_missing; }
const char* kc_fasta_class_name(const kc_fasta* f, uint32_t c) {
  return c < f->class_names.size() ? f->class_names[c].c_str() : "";
}
const char* kc_fasta_id(const kc_fasta* f, uint64_t p) { return p < f->n ? f->ids[p].c_str() : ""; }

// ---- DIAMOND hand-off (src/graph/mod.rs:202-220 directories, :253-261 and :273-280 the two one-record
// FASTA files of a kept pair, :304-317 blastp_output.tsv).  The `diamond` subprocesses themselves
// (:266-270, :283-293) are out of scope; what is written here is exactly what they would be started on.
static bool write_file(const std::string& path, const std::string& text) {
  FILE* fh = std::fopen(path.c_str(), "wb");
  if (!fh) return false;
  const bool ok = std::fwrite(text.data(), 1, text.size(), fh) == text.size();
  return std::fclose(fh) == 0 && ok;
}

const char* kc_blastp_header(void) {
  return "query id\tquery length\tsubject id\tsubject length\tquery alignment start\tquery alignment end\t"
         "subject alignment start\tsubject alignment end\talignment length\tpercent identity\tevalue\tbit score\n";
}

int kc_write_handoff(const kc_fasta* f, const kc_edge* edges, uint64_t n_edges, const char* dir,
                     uint64_t* n_files_out) {
  if (!f || (n_edges && !edges) || !dir) return KC_EINVAL;
  const std::string base = std::string(dir);
  // (the reference removes and re-creates both directories, src/graph/mod.rs:202-220)
  for (const char* sub : {"/fasta_files", "/db_files"}) {
    const std::string d = base + sub;
    if (mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) return KC_EINVAL;
  }
  uint64_t n_files = 0;
  auto accession = [](const std::string& id) {  // split_once('|').0 (the reference panics without a '|')
    const size_t bar = id.find('|');
    return bar == std::string::npos ? id : id.substr(0, bar);
  };
  for (uint64_t i = 0; i < n_edges; ++i) {
    const uint32_t ends[2] = {edges[i].a, edges[i].b};  // [reference, query] = vertices_key order
    for (int side = 0; side < 2; ++side) {
      const uint64_t p = ends[side];
      if (p >= f->n) return KC_EINVAL;
      const std::string id = f->ids[p];
      const std::string seq(reinterpret_cast<const char*>(f->residues + f->offsets[p]),
                            (size_t)(f->offsets[p + 1] - f->offsets[p]));
      // format!("fasta_files/{}_{}.fasta", edge_key, accession); format!(">{}\n{}", id, seq)
      const std::string name = base + "/fasta_files/" + std::to_string(i) + "_" + accession(id) + ".fasta";
      if (!write_file(name, ">" + id + "\n" + seq)) return KC_EINVAL;
      ++n_files;
    }
  }
  if (!write_file(base + "/blastp_output.tsv", kc_blastp_header())) return KC_EINVAL;
  if (n_files_out) *n_files_out = n_files;
  return KC_OK;
}

}  // extern "C"
