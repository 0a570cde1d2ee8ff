// host_qc.cpp -- host-side f64 arithmetic on the integer counts produced on the GPU.
// Keeping these decisions in f64 on the host, in the reference's exact expression order, is what
// makes the retained-SNP masks bit-exact.  Follows:
//   QC ladder + mean/sigma  src/prepare.rs:1281-1375
//   HWE chi-square          src/prepare.rs:1641-1745
//   VCF MAF filter          src/vcf.rs:244-266
//   LD-block mapping        src/prepare.rs:1424-1563
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/gpca.h"
#include "host_qc.h"
#include "parallel_for.h"

extern "C" double gpca_hwe_chi_squared_p_value(uint64_t hom1, uint64_t het, uint64_t hom2) {
  const uint64_t total = hom1 + het + hom2;
  if (total == 0) return 1.0;
  const double c1 = 2.0 * (double)hom1 + (double)het;
  const double c2 = 2.0 * (double)hom2 + (double)het;
  const double tot = c1 + c2;
  if (tot <= 1e-9) return 1.0;
  const double f1 = c1 / tot, f2 = c2 / tot;
  const double FREQ_EPSILON = 1e-9;
  if (f1 < FREQ_EPSILON || f2 < FREQ_EPSILON) return 1.0;
  if (std::fabs(f1 + f2 - 1.0) > 1e-6) return 1.0;
  const double e1 = f1 * f1 * (double)total;
  const double eh = 2.0 * f1 * f2 * (double)total;
  const double e2 = f2 * f2 * (double)total;
  const double MIN_E = 1e-9;
  double chi = 0.0;
  if (e1 > MIN_E) {
    const double d = (double)hom1 - e1;
    chi += d * d / e1;
  } else if ((double)hom1 > MIN_E) {
    chi = INFINITY;
  }
  if (std::isfinite(chi)) {
    if (eh > MIN_E) {
      const double d = (double)het - eh;
      chi += d * d / eh;
    } else if ((double)het > MIN_E) {
      chi = INFINITY;
    }
  }
  if (std::isfinite(chi)) {
    if (e2 > MIN_E) {
      const double d = (double)hom2 - e2;
      chi += d * d / e2;
    } else if ((double)hom2 > MIN_E) {
      chi = INFINITY;
    }
  }
  if (std::isnan(chi)) return 1.0;
  if (std::isinf(chi)) return 0.0;
  // statrs ChiSquared(1).cdf(x) = P(1/2, x/2) = erf(sqrt(x/2))
  const double cdf = (chi <= 0.0) ? 0.0 : std::erf(std::sqrt(0.5 * chi));
  if (std::isnan(cdf)) return 1.0;
  return std::max(1.0 - cdf, 0.0);
}

void host_snp_qc(uint64_t n_samples, uint64_t M, const uint32_t* counts /*[M][4] nvalid,n0,n1,n2*/,
                 const gpca_qc_cfg& cfg, uint8_t* keep, float* mean, float* sd, uint8_t* fail_code) {
  parallel_for(M, [&](uint64_t j_lo, uint64_t j_hi) {
  for (uint64_t j = j_lo; j < j_hi; ++j) {
    const uint32_t nv = counts[4 * j + 0], n0 = counts[4 * j + 1], n1 = counts[4 * j + 2], n2 = counts[4 * j + 3];
    uint8_t code = 0;
    float m32 = 0.f, s32 = 0.f;
    do {
      const double call_rate = (double)nv / (double)n_samples;  // :1283
      if (call_rate < cfg.min_call_rate) { code = 1; break; }
      if (nv == 0) { code = 2; break; }                         // :1292
      const double dsum = (double)n1 + 2.0 * (double)n2;        // exact integer in f64
      const double m = dsum / (double)nv;                       // :1294
      const double freq = m / 2.0;                              // :1295
      const double maf = std::min(freq, 1.0 - freq);            // :1296
      if (maf < cfg.min_maf) { code = 3; break; }               // :1299
      if (std::fabs(freq) < 1e-9 || std::fabs(1.0 - freq) < 1e-9) { code = 4; break; }  // :1302
      if (cfg.max_hwe_p < 1.0) {                                // :1306
        const double p = gpca_hwe_chi_squared_p_value(n0, n1, n2);
        if (p <= cfg.max_hwe_p) { code = 5; break; }            // :1310
      }
      // pass 2 (:1316-1352) in closed form: sum over valid calls of (x-mean)^2
      const double d0 = 0.0 - m, d1 = 1.0 - m, d2 = 2.0 - m;
      const double ssd = (double)n0 * (d0 * d0) + (double)n1 * (d1 * d1) + (double)n2 * (d2 * d2);
      const double var = (nv >= 2) ? ssd / (double)(nv - 1) : 0.0;  // :1357-1361
      if (var <= 1e-9) { code = 6; break; }                     // :1363
      m32 = (float)m;                                           // :1313
      s32 = (float)std::sqrt(var);                              // :1364
    } while (0);
    keep[j] = (code == 0);
    if (mean) mean[j] = m32;
    if (sd) sd[j] = s32;
    if (fail_code) fail_code[j] = code;
  }
  });
}

void host_vcf_maf(uint64_t n_samples, uint64_t M, const uint32_t* counts, double maf_threshold, uint8_t* keep,
                  float* mean, float* sd) {
  parallel_for(M, [&](uint64_t j_lo, uint64_t j_hi) {
  for (uint64_t j = j_lo; j < j_hi; ++j) {
    const uint32_t nv = counts[4 * j + 0], n0 = counts[4 * j + 1], n1 = counts[4 * j + 2], n2 = counts[4 * j + 3];
    bool k = false;
    float m32 = 0.f, s32 = 0.f;
    if (nv == n_samples && n_samples > 0) {                     // vcf.rs:227-242: any missing GT drops the variant
      const uint32_t allele_sum = n1 + 2u * n2;                 // vcf.rs:244
      const uint32_t total = (uint32_t)(n_samples * 2);         // vcf.rs:245
      const double p = (double)allele_sum / (double)total;      // vcf.rs:254
      const double maf = std::min(p, 1.0 - p);                  // vcf.rs:255
      if (!(maf < maf_threshold)) {                             // vcf.rs:259
        k = true;
        const double m = (double)allele_sum / (double)n_samples;
        const double d0 = 0.0 - m, d1 = 1.0 - m, d2 = 2.0 - m;
        const double ssd = (double)n0 * (d0 * d0) + (double)n1 * (d1 * d1) + (double)n2 * (d2 * d2);
        double s = (n_samples >= 2) ? std::sqrt(ssd / (double)(n_samples - 1)) : 0.0;
        if (!(s > 1e-9) || !std::isfinite(s)) s = 1.0;          // rfit scale convention
        m32 = (float)m;
        s32 = (float)s;
      }
    }
    keep[j] = k;
    if (mean) mean[j] = m32;
    if (sd) sd[j] = s32;
  }
  });
}

extern "C" int gpca_map_snps_to_ld_blocks(const char* const* snp_chrom, const int32_t* snp_bp, uint64_t n_qc,
                                          const char* const* blk_chrom, const int32_t* blk_start,
                                          const int32_t* blk_end, uint64_t n_blocks, int64_t* pca_pos,
                                          int64_t* block_of, uint64_t* n_pca, uint64_t* n_blocks_out,
                                          uint64_t* sorted_block_order) {
  if ((n_qc && (!snp_chrom || !snp_bp || !pca_pos || !block_of)) || (n_blocks && (!blk_chrom || !blk_start || !blk_end)))
    return GPCA_ERR_INVALID;
  // tag = "{chr}:{start}-{end}" (prepare.rs:1597); blocks with the same tag merge (HashMap key, :1444)
  std::vector<std::string> tags(n_blocks);
  for (uint64_t b = 0; b < n_blocks; ++b)
    tags[b] = std::string(blk_chrom[b]) + ":" + std::to_string(blk_start[b]) + "-" + std::to_string(blk_end[b]);
  // chromosome -> blocks in file order (the reference scans all blocks linearly, :1450; the first
  // match in file order wins, which a per-chromosome list preserves)
  std::map<std::string, std::vector<uint64_t>> by_chr;
  for (uint64_t b = 0; b < n_blocks; ++b) by_chr[blk_chrom[b]].push_back(b);
  std::vector<int64_t> first_block(n_qc, -1);
  uint64_t npca = 0;
  for (uint64_t i = 0; i < n_qc; ++i) {
    pca_pos[i] = -1;
    auto it = by_chr.find(snp_chrom[i]);
    if (it == by_chr.end()) continue;
    for (uint64_t b : it->second) {
      if (snp_bp[i] >= blk_start[b] && snp_bp[i] <= blk_end[b]) {  // :1452-1453 inclusive
        first_block[i] = (int64_t)b;
        break;
      }
    }
    if (first_block[i] >= 0) pca_pos[i] = (int64_t)npca++;        // inputs are in increasing original index (:1467)
  }
  // non-empty tags sorted by string (:1548)
  std::map<std::string, int64_t> tag_rank;
  for (uint64_t i = 0; i < n_qc; ++i)
    if (first_block[i] >= 0) tag_rank[tags[first_block[i]]] = 0;
  int64_t r = 0;
  for (auto& kv : tag_rank) kv.second = r++;
  if (sorted_block_order) {
    // representative input block index for each sorted tag (first block in file order with that tag)
    std::vector<int64_t> rep(tag_rank.size(), -1);
    for (uint64_t b = 0; b < n_blocks; ++b) {
      auto it = tag_rank.find(tags[b]);
      if (it != tag_rank.end() && rep[it->second] < 0) rep[it->second] = (int64_t)b;
    }
    for (size_t k = 0; k < rep.size(); ++k) sorted_block_order[k] = (uint64_t)rep[k];
  }
  for (uint64_t i = 0; i < n_qc; ++i) block_of[i] = (first_block[i] >= 0) ? tag_rank[tags[first_block[i]]] : -1;
  if (n_pca) *n_pca = npca;
  if (n_blocks_out) *n_blocks_out = (uint64_t)tag_rank.size();
  return GPCA_OK;
}
