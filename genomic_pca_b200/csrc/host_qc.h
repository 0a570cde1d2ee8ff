// host_qc.h -- host-side QC arithmetic (see host_qc.cpp)
#pragma once
#include <stdint.h>
#include "../../include/gpca.h"

void host_snp_qc(uint64_t n_samples, uint64_t M, const uint32_t* counts, const gpca_qc_cfg& cfg, uint8_t* keep,
                 float* mean, float* sd, uint8_t* fail_code);
void host_vcf_maf(uint64_t n_samples, uint64_t M, const uint32_t* counts, double maf_threshold, uint8_t* keep,
                  float* mean, float* sd);
