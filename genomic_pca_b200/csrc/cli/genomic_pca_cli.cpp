// genomic_pca_cli.cpp -- command-line front end with the reference's flag surface and output files,
// driving the hot path through the C ABI (include/gpca.h).  No CPU fallback: gpca_init fails without a B200.
//
// Mirrors (names, defaults, messages, output formats):
//   CliArgs                         src/main.rs:505-592   (effective defaults via default_value_if: main.rs:545-588)
//   run_vcf_workflow                src/main.rs:133-247   (file discovery :139-152, sample set from the first file :157-160)
//   process_single_vcf              src/vcf.rs:65-287     (biallelic single-base filter, GT -> 0/1/2, drop on missing, MAF)
//   run_eigensnp_rust_workflow      src/main.rs:250-447
//   MicroarrayDataPreparer          src/prepare.rs:922-1096 (BIM/FAM metadata, sample keep list), :1565-1616 (LD file)
//   output_writer                   src/main.rs:682-839   (file names, headers, {:.6})
// Deliberate differences: VCF text (plain, gzip or BGZF) is inflated with zlib and parsed here (the reference uses
// noodles-vcf on one thread per file): BGZF blocks are inflated concurrently and the lines of a batch are parsed on all
// host threads straight into 2-bit rows (no D x N byte matrix on the host);
// --threads / --log-level are accepted; the rfit eigenvalues file is header-only exactly like the reference
// (src/main.rs:676) unless --write-eigenvalues is given.
#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../../include/gpca.h"

namespace {

struct Args {
  std::string out;
  std::string vcf_dir, bed_file, ld_block_file, keep_file, log_level = "Info";
  long threads = -1;
  long components = -1;
  double maf = -1.0;
  bool has_seed = false;
  uint64_t rfit_seed = 0;
  bool eigensnp = false;
  bool write_eigenvalues = false;
  uint32_t rfit_power_iters = 2;
  long gpus = 1;            // (extension) GPUs to shard the SNPs / LD blocks over; 0 = every visible one
  gpca_qc_cfg qc{0.98, 0.01, 1e-6};
  gpca_eigensnp_cfg es;
};

[[noreturn]] void die(const std::string& msg) {
  fprintf(stderr, "Error: %s\n", msg.c_str());
  exit(1);
}
void info(const std::string& msg) { fprintf(stderr, "[INFO] %s\n", msg.c_str()); }
void warn(const std::string& msg) { fprintf(stderr, "[WARN] %s\n", msg.c_str()); }

void usage() {
  fprintf(stderr,
          "Genomic PCA Tool from VCF or BED/LD-block files (B200-native hot path).\n\n"
          "Usage: genomic_pca --out <OUTPUT_PREFIX> [OPTIONS]\n\n"
          "  -o, --out <PREFIX>            Output file prefix (required)\n"
          "  -t, --threads <N>             host threads for the host-side stages (default: all cores, split over the GPUs)\n"
          "      --log-level <LEVEL>       accepted for compatibility\n"
          "  -d, --vcf-dir <DIR>           Directory containing VCF files (required unless --eigensnp)\n"
          "  -k, --components <K>          Number of principal components (VCF workflow)\n"
          "      --maf <MAF>               Minimum MAF for VCF variant filtering (default 0.01)\n"
          "      --rfit-seed <SEED>        Seed for the randomized SVD (VCF workflow)\n"
          "      --eigensnp                EigenSNP workflow (requires --bed-file and --ld-block-file)\n"
          "      --bed-file <PATH>         PLINK .bed (with .bim/.fam beside it)\n"
          "      --ld-block-file <PATH>    LD block definitions: chr start end\n"
          "      --eigensnp-sample-keep-file <PATH>\n"
          "      --eigensnp-min-call-rate <F> (0.98)  --eigensnp-min-maf <F> (0.01)  --eigensnp-max-hwe-p <F> (1e-6)\n"
          "      --eigensnp-k-global <N> (10)  --eigensnp-components-per-block <N> (7)\n"
          "      --eigensnp-subset-factor <F> (0.075)  --eigensnp-min-subset-size <N> (10000)\n"
          "      --eigensnp-max-subset-size <N> (40000)  --eigensnp-global-oversampling <N> (10)\n"
          "      --eigensnp-global-power-iter <N> (2)  --eigensnp-local-oversampling <N> (10)\n"
          "      --eigensnp-local-power-iter <N> (2)  --eigensnp-seed <N> (2025)\n"
          "      --eigensnp-snp-strip-size <N> (2000)  --eigensnp-refine-passes <N> (1)\n"
          "      --eigensnp-collect-diagnostics\n"
          "      --write-eigenvalues       (extension) also write the rfit eigenvalues\n"
          "      --gpus <N>                (extension) shard the SNPs / whole LD blocks over N GPUs, NCCL exchange of the\n"
          "                                sample-side sketch (0 = every visible GPU; default 1)\n");
}

// Rust `{:.6}` (src/main.rs:721,755,779,832) = the exact decimal expansion rounded half-to-even to six places, which
// is what glibc's "%.6f" prints -- except NaN, which Rust spells "NaN" (no sign) where C prints "nan" / "-nan".
// Negative zero keeps its sign and the infinities are "inf" / "-inf" in both.
// Fast path for |v| < 2^52 / 1e6: v * 1e6 = p + e exactly with p the rounded product and e = fma(v, 1e6, -p) (1e6 is an
// exact double), so rounding p + e half-to-even to an integer is decided without error; the digits are then printed
// from the integer.  Everything else (huge values, infinities) goes through snprintf.  tests/test_cli.py compares the
// output with an exact decimal expansion on random values and exact ties.
inline int format_f6(char* out, size_t cap, double v) {
  if (v != v) return snprintf(out, cap, "NaN");
  const double av = std::fabs(v);
  if (!(av < 4.0e9) || cap < 24) return snprintf(out, cap, "%.6f", v);
  const double p = av * 1e6;
  const double e = std::fma(av, 1e6, -p);
  double n = std::nearbyint(p);                 // ties-to-even on p (default rounding mode)
  const double d = (p - n) + e;                 // exact: |p - n| <= 0.5 and e is tiny
  const bool odd = std::fmod(n, 2.0) != 0.0;
  if (d > 0.5 || (d == 0.5 && odd)) n += 1.0;
  else if (d < -0.5 || (d == -0.5 && odd)) n -= 1.0;
  uint64_t u = (uint64_t)n;
  const uint64_t ip = u / 1000000u;
  uint32_t fp = (uint32_t)(u % 1000000u);
  char* q = out;
  if (std::signbit(v)) *q++ = '-';
  char tmp[24];
  int len = 0;
  uint64_t t = ip;
  do {
    tmp[len++] = (char)('0' + t % 10);
    t /= 10;
  } while (t);
  while (len) *q++ = tmp[--len];
  *q++ = '.';
  for (int i = 5; i >= 0; --i) {
    q[i] = (char)('0' + fp % 10);
    fp /= 10;
  }
  q += 6;
  *q = 0;
  return (int)(q - out);
}
Args parse(int argc, char** argv) {
  Args a;
  gpca_eigensnp_default_cfg(&a.es);
  auto need = [&](int& i) -> std::string {
    if (i + 1 >= argc) die(std::string("missing value for ") + argv[i]);
    return argv[++i];
  };
  for (int i = 1; i < argc; ++i) {
    std::string f = argv[i];
    std::string inline_val;
    const size_t eq = f.find('=');
    bool has_inline = false;
    if (f.rfind("--", 0) == 0 && eq != std::string::npos) {
      inline_val = f.substr(eq + 1);
      f = f.substr(0, eq);
      has_inline = true;
    }
    auto val = [&]() -> std::string { return has_inline ? inline_val : need(i); };
    if (f == "-h" || f == "--help") { usage(); exit(0); }
    else if (f == "--format-f6") {      // (test hook) the writers' number formatting: one value per argument, as f64 and as f32
      for (int j = i + 1; j < argc; ++j) {
        char t64[400], t32[400];
        const double v = strtod(argv[j], nullptr);
        format_f6(t64, sizeof t64, v);
        format_f6(t32, sizeof t32, (double)(float)v);
        printf("%s\t%s\n", t64, t32);
      }
      exit(0);
    }
    else if (f == "-V" || f == "--version") { printf("genomic_pca %s\n", gpca_version()); exit(0); }
    else if (f == "-o" || f == "--out") a.out = val();
    else if (f == "-t" || f == "--threads") a.threads = atol(val().c_str());
    else if (f == "--log-level") a.log_level = val();
    else if (f == "-d" || f == "--vcf-dir") a.vcf_dir = val();
    else if (f == "-k" || f == "--components") a.components = atol(val().c_str());
    else if (f == "--maf") a.maf = atof(val().c_str());
    else if (f == "--rfit-seed") { a.rfit_seed = strtoull(val().c_str(), nullptr, 10); a.has_seed = true; }
    else if (f == "--rfit-power-iter") a.rfit_power_iters = (uint32_t)atol(val().c_str());
    else if (f == "--eigensnp") a.eigensnp = true;
    else if (f == "--bed-file") a.bed_file = val();
    else if (f == "--ld-block-file") a.ld_block_file = val();
    else if (f == "--eigensnp-sample-keep-file") a.keep_file = val();
    else if (f == "--eigensnp-min-call-rate") a.qc.min_call_rate = atof(val().c_str());
    else if (f == "--eigensnp-min-maf") a.qc.min_maf = atof(val().c_str());
    else if (f == "--eigensnp-max-hwe-p") a.qc.max_hwe_p = atof(val().c_str());
    else if (f == "--eigensnp-k-global") a.es.target_num_global_pcs = (uint32_t)atol(val().c_str());
    else if (f == "--eigensnp-components-per-block") a.es.components_per_ld_block = (uint32_t)atol(val().c_str());
    else if (f == "--eigensnp-subset-factor") a.es.subset_factor = atof(val().c_str());
    else if (f == "--eigensnp-min-subset-size") a.es.min_subset_size = strtoull(val().c_str(), nullptr, 10);
    else if (f == "--eigensnp-max-subset-size") a.es.max_subset_size = strtoull(val().c_str(), nullptr, 10);
    else if (f == "--eigensnp-global-oversampling") a.es.global_oversampling = (uint32_t)atol(val().c_str());
    else if (f == "--eigensnp-global-power-iter") a.es.global_power_iters = (uint32_t)atol(val().c_str());
    else if (f == "--eigensnp-local-oversampling") a.es.local_oversampling = (uint32_t)atol(val().c_str());
    else if (f == "--eigensnp-local-power-iter") a.es.local_power_iters = (uint32_t)atol(val().c_str());
    else if (f == "--eigensnp-seed") a.es.random_seed = strtoull(val().c_str(), nullptr, 10);
    else if (f == "--eigensnp-snp-strip-size") a.es.snp_processing_strip_size = (uint32_t)atol(val().c_str());
    else if (f == "--eigensnp-refine-passes") a.es.refine_pass_count = (uint32_t)atol(val().c_str());
    else if (f == "--eigensnp-collect-diagnostics") a.es.collect_diagnostics = 1;
    else if (f == "--write-eigenvalues") a.write_eigenvalues = true;
    else if (f == "--gpus") a.gpus = atol(val().c_str());
    else die("unexpected argument '" + f + "' (see --help)");
  }
  if (a.out.empty()) die("the following required arguments were not provided: --out <OUTPUT_PREFIX>");
  if (!a.eigensnp) {
    if (a.vcf_dir.empty()) die("--vcf-dir is required for the default VCF workflow.");          // main.rs:116
    if (a.components < 0) die("-k/--components is required for the default VCF workflow.");    // main.rs:119
  } else {
    if (a.bed_file.empty()) die("--bed-file is required when --eigensnp is used");              // main.rs:296
    if (a.ld_block_file.empty()) die("--ld-block-file is required when --eigensnp is used");    // main.rs:299
  }
  return a;
}

void make_parent_dirs(const std::string& prefix) {      // main.rs:219-225
  const size_t p = prefix.find_last_of('/');
  if (p == std::string::npos || p == 0) return;
  std::string dir = prefix.substr(0, p), cur;
  std::stringstream ss(dir);
  std::string part;
  if (dir[0] == '/') cur = "/";
  while (std::getline(ss, part, '/')) {
    if (part.empty()) continue;
    cur += part + "/";
    mkdir(cur.c_str(), 0777);
  }
}

FILE* create_output(const std::string& prefix, const std::string& suffix) {
  const std::string fn = prefix + "." + suffix;
  FILE* f = fopen(fn.c_str(), "w");
  if (!f) die("Failed to create output file '" + fn + "'");
  return f;
}

// Rows are formatted by all host threads into per-block buffers (snprintf "%.6f" = the reference's `{:.6}`) and written
// in order: at 1M rows x 20 columns the text formatting, not the file system, is the cost (main.rs:696-839).
template <class RowFn>
void write_rows_parallel(FILE* f, uint64_t n_rows, size_t bytes_per_row_hint, RowFn&& format_row) {
  const uint64_t BLOCK = 1u << 16;     // rows per round of threads x chunk
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  const uint64_t per_round = BLOCK * hw;
  std::vector<std::string> bufs(hw);
  for (uint64_t r0 = 0; r0 < n_rows; r0 += per_round) {
    const uint64_t r1 = std::min<uint64_t>(n_rows, r0 + per_round);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < hw; ++t) {
      const uint64_t lo = r0 + (uint64_t)t * BLOCK, hi = std::min<uint64_t>(r1, lo + BLOCK);
      bufs[t].clear();
      if (lo >= hi) continue;
      th.emplace_back([&, t, lo, hi] {
        std::string& b = bufs[t];
        b.reserve((size_t)(hi - lo) * bytes_per_row_hint);
        for (uint64_t i = lo; i < hi; ++i) format_row(i, b);
      });
    }
    for (auto& x : th) x.join();
    for (unsigned t = 0; t < hw; ++t)
      if (!bufs[t].empty()) fwrite(bufs[t].data(), 1, bufs[t].size(), f);
  }
}

inline void append_f6(std::string& b, double v) {
  char tmp[400];      // (1e308 has 309 integer digits)
  tmp[0] = '\t';
  const int len = format_f6(tmp + 1, sizeof tmp - 1, v);
  b.append(tmp, (size_t)len + 1);
}

template <class T>
void write_pcs(const std::string& prefix, const std::string& suffix, const std::vector<std::string>& names,
               const T* scores, uint64_t n, uint32_t k) {
  if (k == 0) return;
  FILE* f = create_output(prefix, suffix);
  fputs("SampleID", f);
  for (uint32_t j = 1; j <= k; ++j) fprintf(f, "\tPC%u", j);
  fputc('\n', f);
  write_rows_parallel(f, names.size(), 16 + 12 * (size_t)k, [&](uint64_t i, std::string& b) {
    b += names[i];
    for (uint32_t j = 0; j < k; ++j) {
      if (i < n) append_f6(b, (double)scores[i * k + j]);
      else b += "\tNA";
    }
    b += '\n';
  });
  fclose(f);
}

void write_eigenvalues(const std::string& prefix, const std::vector<double>& ev) {
  FILE* f = create_output(prefix, "eigenvalues.tsv");
  fputs("PC\tEigenvalue\n", f);
  for (size_t i = 0; i < ev.size(); ++i) {
    char tmp[400];
    format_f6(tmp, sizeof tmp, ev[i]);
    fprintf(f, "%zu\t%s\n", i + 1, tmp);
  }
  fclose(f);
}

// (measurement hook) the loadings / scores writers at biobank row counts: `genomic_pca --bench-writers ROWS K PATH`
int bench_writers(uint64_t rows, uint32_t k, const std::string& path) {
  std::vector<float> vals(rows * k);
  uint64_t x = 88172645463325252ull;
  for (auto& v : vals) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    v = (float)((double)(x >> 11) / 9007199254740992.0 - 0.5);
  }
  const auto t0 = std::chrono::steady_clock::now();
  FILE* f = fopen(path.c_str(), "w");
  if (!f) die("cannot create " + path);
  fputs("VariantID\tChrom\tPos", f);
  for (uint32_t j = 1; j <= k; ++j) fprintf(f, "\tPC%u_loading", j);
  fputc('\n', f);
  write_rows_parallel(f, rows, 32 + 12 * (size_t)k, [&](uint64_t i, std::string& b) {
    b += "rs";
    b += std::to_string((unsigned long long)i);
    b += "\t1\t";
    b += std::to_string((unsigned long long)(1000 + 10 * i));
    for (uint32_t j = 0; j < k; ++j) append_f6(b, (double)vals[i * k + j]);
    b += '\n';
  });
  fclose(f);
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  struct stat st;
  stat(path.c_str(), &st);
  printf("{\"rows\": %llu, \"k\": %u, \"seconds\": %.3f, \"bytes\": %lld, \"MB_per_s\": %.1f, \"threads\": %u}\n",
         (unsigned long long)rows, k, s, (long long)st.st_size, (double)st.st_size / s / 1e6, std::thread::hardware_concurrency());
  return 0;
}

void check(gpca_ctx* ctx, int rc, const char* what) {
  if (rc != GPCA_OK) die(std::string(what) + " failed: " + gpca_last_error(ctx));
}

// ---- multi-GPU plumbing: one context per GPU, one host thread per context (src/main.rs has nothing to mirror: the
// reference is a single CPU process).  The contexts are created first (device g for g = 0, 1, ... until gpca_init
// refuses), joined into one NCCL communicator by the library, and every stage of a workflow runs as one thread per shard.
struct Shards {
  std::vector<gpca_ctx*> ctx;
  int n() const { return (int)ctx.size(); }
};

Shards open_shards(long want, long threads) {
  Shards s;
  const long cap = want > 0 ? want : 1024;
  for (long g = 0; g < cap; ++g) {
    gpca_ctx* c = nullptr;
    if (gpca_init(&c, (int)g) != GPCA_OK) break;
    s.ctx.push_back(c);
  }
  if (s.ctx.empty()) die("no sm_100 (B200) GPU available: this build has no CPU fallback");
  if (want > 0 && (long)s.ctx.size() < want)
    die("--gpus " + std::to_string(want) + " requested but only " + std::to_string(s.ctx.size()) + " B200 GPU(s) are visible");
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  const unsigned per = (unsigned)std::max<long>(1, (threads > 0 ? threads : (long)hw) / (long)s.ctx.size());
  for (gpca_ctx* c : s.ctx) gpca_set_host_threads(c, per);
  if (s.n() > 1) {
    uint8_t id[GPCA_COMM_ID_BYTES];
    if (gpca_comm_unique_id(id) != GPCA_OK) die("NCCL is not available (libnccl.so.2): cannot use more than one GPU");
    std::vector<std::thread> th;
    std::vector<int> rc(s.n(), 0);
    for (int g = 0; g < s.n(); ++g) th.emplace_back([&, g] { rc[g] = gpca_comm_init(s.ctx[g], id, g, s.n()); });
    for (auto& t : th) t.join();
    for (int g = 0; g < s.n(); ++g) check(s.ctx[g], rc[g], "gpca_comm_init");
    info("Sharding over " + std::to_string(s.n()) + " GPUs (NCCL exchange issued by the library).");
  }
  return s;
}

// fn(g) on one thread per shard; a failure on any shard ends the run with its message
template <class F>
void for_each_shard(const Shards& s, F&& fn) {
  if (s.n() == 1) {
    fn(0);
    return;
  }
  std::vector<std::thread> th;
  for (int g = 0; g < s.n(); ++g) th.emplace_back([&fn, g] { fn(g); });
  for (auto& t : th) t.join();
}

void close_shards(Shards& s) {
  for (gpca_ctx* c : s.ctx) gpca_destroy(c);
  s.ctx.clear();
}

// ---------------------------------------------------------------------------------------------- VCF workflow
// The text of a VCF file as successive batches of whole lines.  Three containers:
//   * BGZF (what bgzip / htslib write, and what the reference's fixtures are: tests/README.md): a concatenation of
//     gzip members of <= 64 KiB, each carrying its own compressed size in a 'BC' extra field.  The block table is read
//     off the mapped file without inflating anything, and the blocks of a batch are inflated CONCURRENTLY on the host
//     threads, each to its own offset (the ISIZE trailers give the offsets by a prefix sum).  CRC32 is verified.
//   * plain gzip: one stream, inflated serially by zlib (gzread).
//   * plain text: read in batches.
// (The reference reads through noodles-bgzf on one thread per FILE, src/main.rs:171-179; a biobank chromosome is one
// file of tens of GB, so the parallelism has to come from inside the file.)
template <class F>
void run_threads(unsigned t, F&& fn) {
  if (t <= 1) {
    fn(0u);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned i = 0; i < t; ++i) th.emplace_back([&fn, i] { fn(i); });
  for (auto& x : th) x.join();
}

struct BgzfBlock {
  size_t data_off, data_len;   // the raw deflate stream inside the mapped file
  uint32_t isize, crc;
};

// parses the gzip member header at p (n bytes available); a BGZF member has FEXTRA with the subfield 'B','C',len 2
bool bgzf_member(const uint8_t* p, size_t n, size_t* total, BgzfBlock* blk, size_t file_off) {
  if (n < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return false;
  const size_t xlen = p[10] | ((size_t)p[11] << 8);
  if (12 + xlen > n) return false;
  size_t q = 12, bsize = 0;
  bool found = false;
  while (q + 4 <= 12 + xlen) {
    const size_t slen = p[q + 2] | ((size_t)p[q + 3] << 8);
    if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) {
      bsize = (p[q + 4] | ((size_t)p[q + 5] << 8)) + 1;
      found = true;
    }
    q += 4 + slen;
  }
  if (!found || bsize < 12 + xlen + 8 || bsize > n) return false;
  if (p[3] & ~4) return false;      // FNAME / FCOMMENT / FHCRC never appear in BGZF members
  *total = bsize;
  blk->data_off = file_off + 12 + xlen;
  blk->data_len = bsize - (12 + xlen) - 8;
  const uint8_t* t = p + bsize - 8;
  blk->crc = t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
  blk->isize = t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
  return true;
}

struct VcfText {
  enum Kind { PLAIN, GZIP, BGZF } kind = PLAIN;
  std::string path;
  int fd = -1;
  const uint8_t* map = nullptr;
  size_t map_len = 0, off = 0;
  gzFile gz = nullptr;
  std::vector<char> carry;     // the unfinished last line of the previous batch
  unsigned threads;
  size_t batch_bytes;
  bool eof = false;
  uint64_t blocks_inflated = 0;

  VcfText(const std::string& p, unsigned t, size_t batch) : path(p), threads(t ? t : 1), batch_bytes(batch) {
    fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) die("cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0) die("cannot stat " + path);
    map_len = (size_t)st.st_size;
    if (map_len) {
      void* m = mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) die("cannot map " + path);
      map = (const uint8_t*)m;
      madvise(m, map_len, MADV_SEQUENTIAL);
    }
    size_t total;
    BgzfBlock b;
    if (map_len >= 2 && map[0] == 0x1f && map[1] == 0x8b) {
      kind = bgzf_member(map, map_len, &total, &b, 0) ? BGZF : GZIP;
      if (kind == GZIP) {
        gz = gzopen(path.c_str(), "rb");
        if (!gz) die("cannot open " + path);
        gzbuffer(gz, 1 << 20);
      }
    }
  }
  ~VcfText() {
    if (gz) gzclose(gz);
    if (map) munmap((void*)map, map_len);
    if (fd >= 0) close(fd);
  }
  VcfText(const VcfText&) = delete;
  VcfText& operator=(const VcfText&) = delete;

  // text <- carry + the next batch, cut after its last newline (the remainder becomes the carry; at the end of the
  // file it is returned as the last line).  false when nothing is left.
  bool next(std::vector<char>& text) {
    if (eof && carry.empty()) return false;
    text.assign(carry.begin(), carry.end());
    carry.clear();
    if (!eof) {
      if (kind == BGZF) fill_bgzf(text);
      else if (kind == GZIP) fill_gzip(text);
      else fill_plain(text);
    }
    if (!eof) {
      size_t cut = text.size();
      while (cut > 0 && text[cut - 1] != '\n') --cut;
      carry.assign(text.begin() + cut, text.end());
      text.resize(cut);
      return true;               // (possibly no complete line yet: the caller just asks again)
    }
    return !text.empty();
  }

  void fill_plain(std::vector<char>& text) {
    const size_t n = std::min(batch_bytes, map_len - off);
    text.insert(text.end(), (const char*)map + off, (const char*)map + off + n);
    off += n;
    eof = off == map_len;
  }
  void fill_gzip(std::vector<char>& text) {
    const size_t base = text.size();
    text.resize(base + batch_bytes);
    size_t got = 0;
    while (got < batch_bytes) {
      const int n = gzread(gz, text.data() + base + got, (unsigned)std::min<size_t>(batch_bytes - got, 1u << 30));
      if (n < 0) die("gzip stream of " + path + " is corrupt");
      if (n == 0) { eof = true; break; }
      got += (size_t)n;
    }
    text.resize(base + got);
  }
  void fill_bgzf(std::vector<char>& text) {
    std::vector<BgzfBlock> blk;
    std::vector<size_t> out_off;
    size_t out = text.size();
    const size_t base = out;
    while (off < map_len && out - base < batch_bytes) {
      size_t total;
      BgzfBlock b;
      if (!bgzf_member(map + off, map_len - off, &total, &b, off))
        die("BGZF block at byte " + std::to_string(off) + " of " + path + " is malformed");
      off += total;
      if (b.isize == 0) continue;            // the empty end-of-file marker (and any empty member)
      blk.push_back(b);
      out_off.push_back(out);
      out += b.isize;
    }
    eof = off >= map_len;
    text.resize(out);
    std::atomic<size_t> next{0};
    std::atomic<int> bad{0};
    run_threads((unsigned)std::min<size_t>(threads, std::max<size_t>(1, blk.size() / 4)), [&](unsigned) {
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
      for (size_t i = next.fetch_add(1); i < blk.size(); i = next.fetch_add(1)) {
        inflateReset(&zs);
        zs.next_in = const_cast<Bytef*>(map + blk[i].data_off);
        zs.avail_in = (uInt)blk[i].data_len;
        zs.next_out = (Bytef*)text.data() + out_off[i];
        zs.avail_out = blk[i].isize;
        const int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.avail_out != 0 ||
            crc32(crc32(0L, Z_NULL, 0), (const Bytef*)text.data() + out_off[i], blk[i].isize) != blk[i].crc)
          bad = 1;
      }
      inflateEnd(&zs);
    });
    if (bad) die("a BGZF block of " + path + " failed to inflate (corrupt data or CRC mismatch)");
    blocks_inflated += blk.size();
  }
};

bool ends_with(const std::string& s, const std::string& suf) {
  return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

// One VCF data line [p, e) -> id + 2-bit row, after the reference's filters; false = the variant is dropped.
//   biallelic single-base REF / one ALT (vcf.rs:109-121); the GT key's position inside FORMAT (vcf.rs:218-224); the first
//   two alleles of every sample, '/' or '|' separated, each exactly "0" or "1" (vcf.rs:52-63, 153-193); any other call
//   drops the variant (vcf.rs:227-242); MAF filter p = sum / 2N in f64 (vcf.rs:244-266).
// The row is written in PLINK coding (dosage 0 -> 11, 1 -> 10, 2 -> 00: what `count_a1` decodes back to the ALT count).
bool parse_variant_line(const char* p, const char* e, size_t n, double maf_thr, uint8_t* row, size_t bps, std::string& id) {
  const char* fld[9];
  size_t flen[9];
  const char* s = p;
  for (int nf = 0; nf < 9; ++nf) {
    const char* t = (const char*)memchr(s, '\t', (size_t)(e - s));
    if (!t) return false;
    fld[nf] = s;
    flen[nf] = (size_t)(t - s);
    s = t + 1;
  }
  if (flen[3] != 1) return false;
  if ((flen[4] == 1 && fld[4][0] == '.') || memchr(fld[4], ',', flen[4])) return false;
  int gt_pos = -1;
  {
    int idx = 0;
    const char* q = fld[8];
    const char* end = fld[8] + flen[8];
    while (q <= end) {
      const char* c = (const char*)memchr(q, ':', (size_t)(end - q));
      const size_t l = c ? (size_t)(c - q) : (size_t)(end - q);
      if (l == 2 && q[0] == 'G' && q[1] == 'T') { gt_pos = idx; break; }
      if (!c) break;
      q = c + 1;
      ++idx;
    }
  }
  if (gt_pos < 0) return false;
  memset(row, 0, bps);
  size_t si = 0;
  uint32_t allele_sum = 0;
  while (si < n && s < e) {
    const char* t = (const char*)memchr(s, '\t', (size_t)(e - s));
    const char* end = t ? t : e;
    const char* q = s;
    for (int g = 0; g < gt_pos && q < end; ++g) {
      const char* c = (const char*)memchr(q, ':', (size_t)(end - q));
      if (!c) { q = end; break; }
      q = c + 1;
    }
    const char* qe = (const char*)memchr(q, ':', (size_t)(end - q));
    if (!qe) qe = end;
    int dos = 0, na = 0;
    const char* r = q;
    while (r < qe && na < 2) {
      const char* sep = r;
      while (sep < qe && *sep != '/' && *sep != '|') ++sep;
      if (sep - r != 1 || (*r != '0' && *r != '1')) return false;
      dos += *r - '0';
      ++na;
      r = sep + 1;
    }
    if (na != 2) return false;
    const uint32_t code = dos == 0 ? 3u : dos == 1 ? 2u : 0u;
    row[si >> 2] |= (uint8_t)(code << (2 * (si & 3)));
    ++si;
    allele_sum += (uint32_t)dos;
    if (!t) break;
    s = t + 1;
  }
  if (si != n || n == 0) return false;
  const double freq = (double)allele_sum / (double)(uint32_t)(n * 2);
  if (std::min(freq, 1.0 - freq) < maf_thr) return false;
  id.assign(fld[0], flen[0]);
  id += ':';
  id.append(fld[1], flen[1]);
  id += ':';
  id += fld[3][0];
  id += ':';
  id.append(fld[4], flen[4]);
  return true;
}

struct VcfSet {
  std::vector<std::string> samples, ids;
  std::vector<uint8_t> packed;          // variant-major rows of ceil(N / 4) bytes, PLINK coding
  uint64_t bgzf_blocks = 0, batches = 0;
};

// The files in sorted-path order (main.rs:152, vcf.rs:293-315), each one read batch by batch; the lines of a batch are
// split over the host threads at line boundaries and every thread appends to its own (ids, rows) piece; the pieces are
// joined in order.  The sample set is the header of the first file; every other file must repeat it (main.rs:157-160).
VcfSet parse_vcf_files(const std::vector<std::string>& files, double maf_thr, unsigned workers, size_t batch_bytes = 64u << 20) {
  VcfSet out;
  if (workers == 0) workers = 1;
  struct Piece {
    std::vector<std::string> ids;
    std::vector<uint8_t> rows;
  };
  std::vector<char> text;
  for (size_t fi = 0; fi < files.size(); ++fi) {
    VcfText in(files[fi], workers, batch_bytes);
    bool in_header = true, have_gt_format = false, have_chrom = false;
    while (in.next(text)) {
      ++out.batches;
      const char* p = text.data();
      const char* const e = p + text.size();
      // header lines (serial: they come first and are few)
      while (in_header && p < e) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        const char* le = nl ? nl : e;
        const char* lt = le;
        if (lt > p && lt[-1] == '\r') --lt;
        if (lt == p) { p = nl ? nl + 1 : e; continue; }
        if (*p != '#') { in_header = false; break; }
        const std::string line(p, lt);
        if (line.rfind("##FORMAT=<ID=GT", 0) == 0) have_gt_format = true;
        if (line.rfind("#CHROM", 0) == 0) {
          std::vector<std::string> hdr;
          std::stringstream ss(line);
          std::string t;
          int col = 0;
          while (std::getline(ss, t, '\t')) if (col++ >= 9) hdr.push_back(t);
          if (fi == 0) {
            if (hdr.empty()) die("VCF header from " + files[0] + " contains no samples.");                    // vcf.rs:32
            out.samples = hdr;
          } else if (hdr != out.samples) {
            die("Sample mismatch in VCF " + files[fi] + ": all VCFs must match the sample set of the first VCF (" + files[0] + ").");
          }
          if (!have_gt_format) die("GT key (FORMAT=GT) not found in FORMAT header for VCF " + files[fi]);     // vcf.rs:93
          have_chrom = true;
        }
        p = nl ? nl + 1 : e;
      }
      if (p >= e) continue;
      if (!have_chrom) {
        if (fi == 0) die("VCF header from " + files[0] + " contains no samples.");
        die("VCF " + files[fi] + " has no #CHROM header line");
      }
      const size_t n = out.samples.size(), bps = (n + 3) / 4;
      const size_t len = (size_t)(e - p);
      const unsigned T = (unsigned)std::min<size_t>(workers, std::max<size_t>(1, len >> 16));
      std::vector<Piece> pieces(T);
      run_threads(T, [&](unsigned t) {
        // piece t = the lines that START in [len t / T, len (t+1) / T)
        auto line_start = [&](size_t pos) -> const char* {
          if (pos == 0) return p;
          if (pos >= len) return e;
          const char* nl = (const char*)memchr(p + pos - 1, '\n', len - pos + 1);
          return nl ? nl + 1 : e;
        };
        const char* q = line_start(len * t / T);
        const char* const qe = line_start(len * (t + 1) / T);
        Piece& pc = pieces[t];
        std::vector<uint8_t> row(bps);
        std::string id;
        while (q < qe) {
          const char* nl = (const char*)memchr(q, '\n', (size_t)(e - q));
          const char* le = nl ? nl : e;
          const char* lt = le;
          if (lt > q && lt[-1] == '\r') --lt;
          if (lt > q && *q != '#' && parse_variant_line(q, lt, n, maf_thr, row.data(), bps, id)) {
            pc.ids.push_back(id);
            pc.rows.insert(pc.rows.end(), row.begin(), row.end());
          }
          q = nl ? nl + 1 : e;
        }
      });
      size_t add = 0;
      for (const Piece& pc : pieces) add += pc.rows.size();
      if (out.packed.capacity() < out.packed.size() + add) out.packed.reserve(std::max(out.packed.size() + add, out.packed.capacity() * 2));
      for (Piece& pc : pieces) {
        out.ids.insert(out.ids.end(), std::make_move_iterator(pc.ids.begin()), std::make_move_iterator(pc.ids.end()));
        out.packed.insert(out.packed.end(), pc.rows.begin(), pc.rows.end());
      }
    }
    out.bgzf_blocks += in.blocks_inflated;
    if (fi == 0 && out.samples.empty()) die("VCF header from " + files[0] + " contains no samples.");           // vcf.rs:32
  }
  return out;
}

// (test / measurement hook) `genomic_pca --parse-vcf DIR MAF THREADS BATCH_BYTES [DUMP]`: the host side of the VCF
// workflow alone -- no GPU involved; prints what was parsed and optionally dumps ids + packed rows.
std::vector<std::string> discover_vcf_files(const std::string& dir);
int parse_vcf_hook(int argc, char** argv) {
  const std::vector<std::string> files = discover_vcf_files(argv[2]);
  const auto t0 = std::chrono::steady_clock::now();
  VcfSet vs = parse_vcf_files(files, strtod(argv[3], nullptr), (unsigned)atol(argv[4]), (size_t)strtoull(argv[5], nullptr, 10));
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  uint64_t bytes = 0;
  for (const auto& f : files) {
    struct stat st;
    if (stat(f.c_str(), &st) == 0) bytes += (uint64_t)st.st_size;
  }
  printf("{\"files\": %zu, \"samples\": %zu, \"variants\": %zu, \"bgzf_blocks\": %llu, \"batches\": %llu, \"threads\": %ld, "
         "\"seconds\": %.3f, \"file_MB_per_s\": %.1f}\n",
         files.size(), vs.samples.size(), vs.ids.size(), (unsigned long long)vs.bgzf_blocks, (unsigned long long)vs.batches,
         atol(argv[4]), s, (double)bytes / s / 1e6);
  if (argc > 6) {
    FILE* f = fopen((std::string(argv[6]) + ".ids").c_str(), "w");
    if (!f) die("cannot create dump");
    for (const auto& id : vs.ids) fprintf(f, "%s\n", id.c_str());
    fclose(f);
    f = fopen((std::string(argv[6]) + ".packed").c_str(), "wb");
    if (!f) die("cannot create dump");
    fwrite(vs.packed.data(), 1, vs.packed.size(), f);
    fclose(f);
  }
  return 0;
}

// discovery: regular files with extension vcf or gz whose name contains ".vcf", sorted (main.rs:139-152)
std::vector<std::string> discover_vcf_files(const std::string& dir) {
  std::vector<std::string> files;
  DIR* d = opendir(dir.c_str());
  if (!d) die("cannot read directory " + dir);
  while (dirent* e = readdir(d)) {
    const std::string name = e->d_name;
    const std::string path = dir + "/" + name;
    struct stat st;
    if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
    if ((ends_with(name, ".vcf") || ends_with(name, ".gz")) && name.find(".vcf") != std::string::npos) files.push_back(path);
  }
  closedir(d);
  if (files.empty()) die("No VCF files (ending in .vcf or .vcf.gz) found in directory: " + dir);
  std::sort(files.begin(), files.end());
  return files;
}

int run_vcf(const Args& a) {
  const std::vector<std::string> files = discover_vcf_files(a.vcf_dir);
  const double maf_thr = a.maf >= 0 ? a.maf : 0.01;                                 // vcf.rs:257
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  const unsigned workers = (unsigned)(a.threads > 0 ? a.threads : (long)hw);
  VcfSet vs = parse_vcf_files(files, maf_thr, workers);
  const std::vector<std::string>& samples = vs.samples;
  const std::vector<std::string>& ids = vs.ids;
  const std::vector<uint8_t>& packed = vs.packed;
  const uint64_t bps = (samples.size() + 3) / 4;
  const uint64_t n = samples.size(), dvar = ids.size();
  if (dvar == 0) die("No variants passed filters across all VCF files. Cannot proceed with PCA.");   // main.rs:198
  info("Aggregated " + std::to_string(dvar) + " variants in total across all VCFs.");
  if (a.components == 0) die("Number of components (-k) must be > 0.");                              // main.rs:607
  if (n < 2) die("PCA requires at least 2 samples, found " + std::to_string(n) + ".");               // main.rs:614
  // Variants are split into contiguous ranges, one per GPU; every shard filters its range, and the rfit passes exchange
  // the N x l sketch (the scores are the same on every shard; shard 0's copy is written).
  Shards sh = open_shards(a.gpus, a.threads);
  const int G = sh.n();
  std::vector<uint64_t> v0(G + 1);
  for (int g = 0; g <= G; ++g) v0[g] = dvar * (uint64_t)g / (uint64_t)G;
  std::vector<uint64_t> kept(G, 0);
  for_each_shard(sh, [&](int g) {
    gpca_ctx* ctx = sh.ctx[g];
    const uint64_t dv = v0[g + 1] - v0[g];
    // 2-bit rows in PLINK coding, packed while parsing: a quarter of the bytes of the u8 matrix on the host and on the bus
    check(ctx, gpca_load_bed(ctx, packed.data() + v0[g] * bps, n, dv, nullptr, 0), "load");
    std::vector<uint8_t> keep(dv);
    std::vector<float> mean(dv), sd(dv);
    check(ctx, gpca_vcf_maf_filter(ctx, maf_thr, keep.data(), mean.data(), sd.data()), "maf filter");
    check(ctx, gpca_set_pca_snps_mask(ctx, keep.data(), mean.data(), sd.data(), &kept[g]), "set_pca_snps");
  });
  uint64_t d_kept = 0;
  std::vector<uint64_t> koff(G, 0);
  for (int g = 0; g < G; ++g) {
    koff[g] = d_kept;
    d_kept += kept[g];
  }
  uint32_t k = (uint32_t)a.components;
  const uint64_t maxk = std::min<uint64_t>(n, d_kept);
  if (k > maxk) {
    warn("Requested k=" + std::to_string(k) + " components exceeds max possible for data; adjusting to " + std::to_string(maxk) + ".");
    k = (uint32_t)maxk;
  }
  std::vector<double> scores(n * k), ev(k);
  uint32_t k_out = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for_each_shard(sh, [&](int g) {
    gpca_ctx* ctx = sh.ctx[g];
    if (G > 1) check(ctx, gpca_set_shard(ctx, koff[g], d_kept), "set_shard");
    std::vector<double> ev_g(g != 0 ? k : 0);
    uint32_t ko = 0;
    // (the scores are the same on every shard: only shard 0's copy is brought to the host)
    check(ctx, gpca_rfit(ctx, k, 10 /* main.rs:636 */, a.rfit_power_iters, a.rfit_seed, a.has_seed ? 1 : 0,
                         g == 0 ? scores.data() : nullptr, g == 0 ? ev.data() : ev_g.data(), nullptr, &ko), "rfit");
    if (g == 0) k_out = ko;
  });
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  info("VCF PCA computation (rfit on " + std::to_string(G) + " GPU(s)) completed in " + std::to_string(ms) + " ms.");
  make_parent_dirs(a.out);
  write_pcs<double>(a.out, "vcf.pca.tsv", samples, scores.data(), n, k_out);
  ev.resize(k_out);
  write_eigenvalues(a.out, a.write_eigenvalues ? ev : std::vector<double>());   // reference writes a header-only file (main.rs:676)
  warn("Loadings output for VCF-based PCA is currently skipped (as in the reference, main.rs:233).");
  close_shards(sh);
  return 0;
}

// ------------------------------------------------------------------------------------------ EigenSNP workflow
std::string normalize_chromosome_name(std::string s) {      // prepare.rs:1610-1616
  for (auto& ch : s) ch = (char)tolower((unsigned char)ch);
  while (s.rfind("chr", 0) == 0) s = s.substr(3);
  return s;
}

std::vector<std::string> split_ws(const std::string& s) {
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string t;
  while (ss >> t) out.push_back(t);
  return out;
}

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

int run_eigensnp(const Args& a) {
  std::string prefix = a.bed_file;
  if (ends_with(prefix, ".bed")) prefix = prefix.substr(0, prefix.size() - 4);
  // FAM: IID = 2nd column; BIM: chrom, sid, cm, bp
  std::vector<std::string> iids, chrom, sid;
  std::vector<int32_t> bp;
  {
    std::ifstream f(prefix + ".fam");
    if (!f) die("Failed to open FAM file '" + prefix + ".fam'");
    std::string line;
    while (std::getline(f, line)) {
      auto p = split_ws(line);
      if (p.size() >= 2) iids.push_back(p[1]);
    }
  }
  {
    std::ifstream f(prefix + ".bim");
    if (!f) die("Failed to open BIM file '" + prefix + ".bim'");
    std::string line;
    while (std::getline(f, line)) {
      auto p = split_ws(line);
      if (p.size() >= 4) {
        chrom.push_back(p[0]);
        sid.push_back(p[1]);
        bp.push_back((int32_t)atol(p[3].c_str()));
      }
    }
  }
  const uint64_t n_fam = iids.size(), m = sid.size();
  info("Initial metadata loaded: " + std::to_string(n_fam) + " samples, " + std::to_string(m) + " SNPs.");
  // (the payload is streamed from the file by gpca_ingest_bed_file below: never held in host memory as a whole)
  // sample QC (prepare.rs:1058-1096)
  std::vector<int64_t> keep_idx;
  bool use_keep = false;
  if (!a.keep_file.empty()) {
    std::ifstream f(a.keep_file);
    if (!f) die("Failed to read sample ID file '" + a.keep_file + "'");
    std::unordered_set<std::string> ks;
    std::string line;
    while (std::getline(f, line)) ks.insert(line);
    for (uint64_t i = 0; i < n_fam; ++i)
      if (ks.count(iids[i])) keep_idx.push_back((int64_t)i);
    use_keep = true;
    if (keep_idx.empty()) die("No samples passed QC.");                                   // prepare.rs:1010
  } else {
    warn("No external sample ID list provided; using all " + std::to_string(n_fam) + " initial samples.");
  }
  const uint64_t n = use_keep ? keep_idx.size() : n_fam;
  // LD block file (prepare.rs:1565-1607)
  std::vector<std::string> bchr;
  std::vector<int32_t> bstart, bend;
  {
    std::ifstream f(a.ld_block_file);
    if (!f) die("Failed to open LD block file '" + a.ld_block_file + "'");
    std::string line;
    while (std::getline(f, line)) {
      const std::string t = trim(line);
      if (t.empty() || t[0] == '#' || t.rfind("chr\t", 0) == 0 || t.rfind("chromosome\t", 0) == 0) continue;
      auto p = split_ws(t);
      if (p.size() < 3) { warn("Skipping malformed LD block line: '" + line + "'"); continue; }
      bchr.push_back(normalize_chromosome_name(p[0]));
      bstart.push_back((int32_t)atol(p[1].c_str()));
      bend.push_back((int32_t)atol(p[2].c_str()));
    }
  }
  std::vector<const char*> bchr_c(bchr.size());
  for (size_t b = 0; b < bchr.size(); ++b) bchr_c[b] = bchr[b].c_str();
  std::vector<std::string> nchr(m);
  for (uint64_t j = 0; j < m; ++j) nchr[j] = normalize_chromosome_name(chrom[j]);
  // SNP -> first matching block (prepare.rs:1447-1463) depends only on chromosome / position, so it is evaluated for
  // EVERY BIM row before the genotypes are read: rows outside every block never reach the resident matrices (the
  // ingest mask), and whole blocks can be dealt to the GPUs.  The numbering of PcaSnpIds and the tag-sorted block
  // lists (prepare.rs:1465-1549) are made from the rows that also pass QC, per shard, after the ingest.
  auto map_rows = [&](const std::vector<uint64_t>& rows, std::vector<int64_t>& pca_pos, std::vector<int64_t>& block_of,
                      uint64_t& n_pca, uint64_t& n_blk) {
    std::vector<const char*> qchr(rows.size());
    std::vector<int32_t> qbp(rows.size());
    for (size_t i = 0; i < rows.size(); ++i) {
      qchr[i] = nchr[rows[i]].c_str();
      qbp[i] = bp[rows[i]];
    }
    pca_pos.assign(rows.size(), -1);
    block_of.assign(rows.size(), -1);
    std::vector<uint64_t> order(std::max<size_t>(bchr.size(), 1));
    if (gpca_map_snps_to_ld_blocks(qchr.data(), qbp.data(), rows.size(), bchr_c.data(), bstart.data(), bend.data(),
                                   bchr.size(), pca_pos.data(), block_of.data(), &n_pca, &n_blk, order.data()) != GPCA_OK)
      die("LD block mapping failed");
  };
  std::vector<uint64_t> all_rows(m);
  for (uint64_t j = 0; j < m; ++j) all_rows[j] = j;
  std::vector<int64_t> pos_all, blk_all;
  uint64_t n_in_blocks = 0, n_blk_all = 0;
  map_rows(all_rows, pos_all, blk_all, n_in_blocks, n_blk_all);
  if (n_in_blocks == 0) die("No SNPs mapped to LD blocks or all resulting blocks were empty.");  // prepare.rs:1031

  Shards sh = open_shards(a.gpus, a.threads);
  const int G = sh.n();
  // whole blocks to GPUs: blocks in the order of their first BIM row, cut where the running SNP count passes g/G
  std::vector<uint64_t> blk_first(n_blk_all, m), blk_count(n_blk_all, 0);
  for (uint64_t j = 0; j < m; ++j)
    if (blk_all[j] >= 0) {
      blk_first[blk_all[j]] = std::min<uint64_t>(blk_first[blk_all[j]], j);
      blk_count[blk_all[j]]++;
    }
  std::vector<uint64_t> by_row(n_blk_all);
  for (uint64_t b = 0; b < n_blk_all; ++b) by_row[b] = b;
  std::sort(by_row.begin(), by_row.end(), [&](uint64_t x, uint64_t y) { return blk_first[x] < blk_first[y]; });
  std::vector<int> shard_of_blk(n_blk_all, 0);
  {
    uint64_t run = 0;
    for (uint64_t q = 0; q < n_blk_all; ++q) {
      const uint64_t b = by_row[q];
      shard_of_blk[b] = (int)std::min<uint64_t>(G - 1, run * (uint64_t)G / n_in_blocks);
      run += blk_count[b];
    }
  }
  struct ShardData {
    uint64_t row0 = 0, row1 = 0, n_blocks = 0;
    std::vector<uint8_t> mask, keep;
    std::vector<float> mean, sd;
    uint64_t n_kept = 0, id_offset = 0;
    std::vector<uint64_t> pca_orig;                 // BIM row of every local PcaSnpId
    std::vector<uint64_t> offs, flat;               // LD blocks, local ids, tag-sorted
    std::vector<float> scores, loadings;
    std::vector<double> ev;
    uint32_t k_out = 0;
  };
  std::vector<ShardData> sd_(G);
  for (int g = 0; g < G; ++g) sd_[g].row0 = m;
  for (uint64_t j = 0; j < m; ++j)
    if (blk_all[j] >= 0) {
      ShardData& d = sd_[shard_of_blk[blk_all[j]]];
      d.row0 = std::min(d.row0, j);
      d.row1 = std::max(d.row1, j + 1);
    }
  for (int g = 0; g < G; ++g) {
    ShardData& d = sd_[g];
    if (d.row1 <= d.row0) die("fewer LD blocks with SNPs than GPUs: use a smaller --gpus");
    const uint64_t nr = d.row1 - d.row0;
    d.mask.assign(nr, 0);
    std::set<int64_t> blks;
    for (uint64_t j = d.row0; j < d.row1; ++j)
      if (blk_all[j] >= 0 && shard_of_blk[blk_all[j]] == g) {
        d.mask[j - d.row0] = 1;
        blks.insert(blk_all[j]);
      }
    d.n_blocks = blks.size();
    d.keep.resize(nr);
    d.mean.resize(nr);
    d.sd.resize(nr);
  }
  // load + SNP QC + resident matrices in one streaming pass over each shard's rows of the file (prepare.rs:995-1098);
  // the memory the EigenSNP driver will need is left free (decides whether both orientations stay resident)
  for_each_shard(sh, [&](int g) {
    gpca_ctx* ctx = sh.ctx[g];
    ShardData& d = sd_[g];
    const uint64_t nr = d.row1 - d.row0;
    check(ctx, gpca_set_memory_reserve(ctx, gpca_eigensnp_workspace_bytes(n, nr, d.n_blocks, &a.es)), "memory reserve");
    check(ctx, gpca_set_ingest_mask(ctx, d.mask.data(), nr), "ingest mask");
    const int rc = gpca_ingest_bed_file_rows(ctx, a.bed_file.c_str(), n_fam, m, d.row0, nr,
                                             use_keep ? keep_idx.data() : nullptr, keep_idx.size(), &a.qc, 0.0,
                                             d.keep.data(), d.mean.data(), d.sd.data(), nullptr, &d.n_kept);
    if (rc != GPCA_OK) die(gpca_last_error(ctx));
  });
  uint64_t n_pca = 0, n_blk = 0;
  for (int g = 0; g < G; ++g) {
    ShardData& d = sd_[g];
    d.id_offset = n_pca;
    std::vector<uint64_t> rows;
    for (uint64_t t = 0; t < d.keep.size(); ++t)
      if (d.keep[t]) rows.push_back(d.row0 + t);
    std::vector<int64_t> pca_pos, block_of;
    uint64_t np = 0, nb = 0;
    map_rows(rows, pca_pos, block_of, np, nb);
    if (np != rows.size()) die("internal: a SNP that passed the ingest mask lies in no LD block");
    std::vector<std::vector<uint64_t>> blocks(nb);
    for (size_t i = 0; i < rows.size(); ++i) blocks[block_of[i]].push_back((uint64_t)pca_pos[i]);
    d.pca_orig = rows;
    d.offs.assign(nb + 1, 0);
    for (uint64_t b = 0; b < nb; ++b) {
      d.offs[b + 1] = d.offs[b] + blocks[b].size();
      d.flat.insert(d.flat.end(), blocks[b].begin(), blocks[b].end());
    }
    n_pca += np;
    n_blk += nb;
  }
  info("SNP QC & Stats calculation complete. " + std::to_string(n_pca) + " / " + std::to_string(m) +
       " initial SNPs passed all filters and lie in an LD block.");
  if (n_pca == 0) die("No SNPs passed all QC filters.");                                  // prepare.rs:1020
  info("LD Mapping: " + std::to_string(n_pca) + " unique SNPs (D_blocked) mapped to " + std::to_string(n_blk) + " LD blocks.");
  const uint32_t k = a.es.target_num_global_pcs;
  const auto t0 = std::chrono::steady_clock::now();
  for_each_shard(sh, [&](int g) {
    gpca_ctx* ctx = sh.ctx[g];
    ShardData& d = sd_[g];
    if (G > 1) check(ctx, gpca_set_shard(ctx, d.id_offset, n_pca), "set_shard");
    if (g == 0) d.scores.resize(n * k);       // (the same on every shard: only shard 0's copy is brought to the host)
    d.loadings.resize(d.n_kept * k);
    d.ev.resize(k);
    check(ctx, gpca_eigensnp(ctx, &a.es, d.offs.data(), d.offs.size() - 1, d.flat.data(), g == 0 ? d.scores.data() : nullptr, d.ev.data(),
                             d.loadings.data(), &d.k_out), "eigensnp");
  });
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  info("EigenSNP PCA algorithm completed in " + std::to_string(ms) + " ms.");
  const uint32_t k_out = sd_[0].k_out;
  // outputs (main.rs:382-408); the scores and eigenvalues are the same on every shard
  make_parent_dirs(a.out);
  std::vector<std::string> names;
  if (use_keep) for (int64_t i : keep_idx) names.push_back(iids[i]);
  else names = iids;
  write_pcs<float>(a.out, "eigensnp.pca.tsv", names, sd_[0].scores.data(), n, k_out);
  std::vector<double> ev(sd_[0].ev.begin(), sd_[0].ev.begin() + k_out);
  write_eigenvalues(a.out, ev);
  if (k_out > 0) {
    // loadings rows in increasing BIM index (PcaSnpId order, prepare.rs:1465-1486): the shards' row ranges are merged
    struct Ref { uint64_t orig; int g; uint64_t local; };
    std::vector<Ref> refs;
    refs.reserve(n_pca);
    for (int g = 0; g < G; ++g)
      for (uint64_t i = 0; i < sd_[g].pca_orig.size(); ++i) refs.push_back({sd_[g].pca_orig[i], g, i});
    if (!std::is_sorted(refs.begin(), refs.end(), [](const Ref& x, const Ref& y) { return x.orig < y.orig; }))
      std::sort(refs.begin(), refs.end(), [](const Ref& x, const Ref& y) { return x.orig < y.orig; });
    FILE* f = create_output(a.out, "eigensnp.loadings.tsv");
    fputs("VariantID\tChrom\tPos", f);
    for (uint32_t j = 1; j <= k_out; ++j) fprintf(f, "\tPC%u_loading", j);
    fputc('\n', f);
    write_rows_parallel(f, n_pca, 32 + 12 * (size_t)k_out, [&](uint64_t i, std::string& b) {
      const Ref& r = refs[i];
      const uint64_t o = r.orig;
      b += sid[o];
      b += '\t';
      b += chrom[o];
      b += '\t';
      b += std::to_string((unsigned long long)(uint64_t)bp[o]);
      const float* row = sd_[r.g].loadings.data() + r.local * k_out;
      for (uint32_t j = 0; j < k_out; ++j) append_f6(b, (double)row[j]);
      b += '\n';
    });
    fclose(f);
  }
  if (a.es.collect_diagnostics) {      // main.rs:411-430
    const std::string fn = a.out + ".eigensnp_diagnostics.json";
    info("Writing EigenSNP diagnostics to " + fn + "...");
    FILE* f = fopen(fn.c_str(), "w");
    if (!f) {
      warn("Failed to write EigenSNP diagnostics to " + fn);
    } else {
      if (G == 1) {
        fputs(gpca_eigensnp_diagnostics(sh.ctx[0]), f);
      } else {
        fputs("{\"shards\": [\n", f);
        for (int g = 0; g < G; ++g) {
          fputs(gpca_eigensnp_diagnostics(sh.ctx[g]), f);
          if (g + 1 < G) fputs(",\n", f);
        }
        fputs("]}\n", f);
      }
      fclose(f);
    }
  }
  close_shards(sh);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc == 5 && std::string(argv[1]) == "--bench-writers")
    return bench_writers(strtoull(argv[2], nullptr, 10), (uint32_t)atol(argv[3]), argv[4]);
  if (argc >= 6 && std::string(argv[1]) == "--parse-vcf") return parse_vcf_hook(argc, argv);
  const Args a = parse(argc, argv);
  const auto t0 = std::chrono::steady_clock::now();
  const int rc = a.eigensnp ? run_eigensnp(a) : run_vcf(a);
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  info(std::string("genomic_pca (mode: ") + (a.eigensnp ? "EigenSNP" : "VCF") + ") finished successfully in " + std::to_string(s) + " s.");
  return rc;
}
