// Host-only part of the C ABI: option defaults, error slot, synthetic fixtures.
#include "../../include/fastquick_b200.h"
#include "fq_index.h"
#include "fq_synth.h"
#include "fq_common.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <zlib.h>

namespace fqb {
static thread_local std::string g_err;
void set_error(const std::string &e) { g_err = e; }
}  // namespace fqb

extern "C" {

const char *fqb_last_error(void) { return fqb::g_err.c_str(); }

// 1-based number of the drand48() call that returns exactly 0.0 after srand48(seed).  X(n+1) = (a X(n) + c) mod 2^48 with a = 1
// mod 4 and c odd has full period modulo every power of two, so X(n) mod 2^k depends on n mod 2^k only and the n with X(n) = 0
// is found one bit at a time (48 jump-aheads).
uint64_t fqb_drand48_zero_index(uint32_t seed) {
    const uint64_t a0 = 0x5DEECE66Dull, c0 = 0xBull, mask = 0xFFFFFFFFFFFFull;
    const uint64_t x0 = ((uint64_t)seed << 16) | 0x330Eull;              // srand48
    auto advance = [&](uint64_t n) {                                      // state after n calls
        uint64_t a = a0, c = c0, A = 1, Cc = 0;
        while (n) {
            if (n & 1) { A = (A * a) & mask; Cc = (Cc * a + c) & mask; }
            c = ((a + 1) * c) & mask; a = (a * a) & mask; n >>= 1;
        }
        return (A * x0 + Cc) & mask;
    };
    uint64_t n = 0;
    for (int k = 0; k < 48; ++k)
        if (advance(n) & ((2ull << k) - 1)) n |= 1ull << k;
    return n ? n : 1ull << 48;                                            // n == 0: the seed state itself, met again after a full period
}

// gap_init_opt(), libbwa/bwtaln.c:24-48; kmer_thresh: src/FASTQuick.cpp:174
void fqb_gap_opt_default(fqb_gap_opt_t *o) {
    memset(o, 0, sizeof(*o));
    o->s_mm = 3; o->s_gapo = 11; o->s_gape = 4;
    o->max_diff = -1; o->max_gapo = 1; o->max_gape = 6;
    o->indel_end_skip = 5; o->max_del_occ = 10; o->max_entries = 2000000;
    o->mode = 0x01 | 0x02;               /* BWA_MODE_GAPE | BWA_MODE_COMPREAD */
    o->seed_len = 32; o->max_seed_diff = 2;
    o->fnr = 0.02;
    o->max_top2 = 30;
    o->trim_qual = 0;
    o->flank_len = 250; o->flank_long_len = 1000;
    o->read_len = 151;
    o->kmer_thresh = 3;
    o->is_il13 = 0;
}
// bwa_init_pe_opt(), libbwa/bwape.c:7-20
void fqb_pe_opt_default(fqb_pe_opt_t *o) {
    memset(o, 0, sizeof(*o));
    o->max_isize = 500; o->force_isize = 0; o->max_occ = 100000;
    o->n_multi = 3; o->N_multi = 10; o->type = 1; o->is_sw = 1; o->ap_prior = 1e-5;
}

struct fqb_synth { fqb::SynthRef ref; };

void fqb_synth_ref_cfg_default(fqb_synth_ref_cfg_t *c) {
    fqb::SynthRefConfig d;
    c->seed = d.seed; c->n_long = d.n_long; c->n_short = d.n_short; c->n_x = d.n_x; c->n_y = d.n_y;
    c->flank_short = d.flank_short; c->flank_long = d.flank_long; c->spacing = d.spacing; c->n_dup = d.n_dup;
}
void fqb_synth_read_cfg_default(fqb_synth_read_cfg_t *c) {
    fqb::SynthReadConfig d;
    c->seed = d.seed; c->read_len = d.read_len; c->max_indel_len = d.max_indel_len; c->f_on = d.f_on;
    c->sub_rate = d.sub_rate; c->ins_rate = d.ins_rate; c->del_rate = d.del_rate; c->n_rate = d.n_rate;
    c->isize_mean = d.isize_mean; c->isize_sd = d.isize_sd; c->bad_tail_rate = d.bad_tail_rate;
}
int fqb_synth_create(const fqb_synth_ref_cfg_t *c, fqb_synth **out) {
    if (!c || !out) { fqb::set_error("null argument"); return FQB_ERR_ARG; }
    fqb::SynthRefConfig cfg;
    cfg.seed = c->seed; cfg.n_long = c->n_long; cfg.n_short = c->n_short; cfg.n_x = c->n_x; cfg.n_y = c->n_y;
    cfg.flank_short = c->flank_short; cfg.flank_long = c->flank_long; cfg.spacing = c->spacing; cfg.n_dup = c->n_dup;
    if (cfg.spacing < 2 * cfg.flank_long + 102 || cfg.n_long + cfg.n_short < 1 || cfg.n_dup < 0 || cfg.flank_short < 230) { fqb::set_error("bad synthetic reference shape"); return FQB_ERR_ARG; }
    fqb_synth *s = new fqb_synth();
    fqb::synth_reference(cfg, s->ref);
    *out = s;
    return FQB_OK;
}
void fqb_synth_destroy(fqb_synth *s) { delete s; }

int fqb_synth_write_inputs(const fqb_synth *s, const char *dir) {
    std::string err;
    if (!fqb::synth_write_reference_inputs(s->ref, dir, err)) { fqb::set_error(err); return FQB_ERR_IO; }
    return FQB_OK;
}
int fqb_synth_write_index(const fqb_synth *s, const char *genome_path, const char *dbsnp_path, const char *prefix, int with_rollhash) {
    std::string err;
    fqb::HostIndex idx;
    fqb::build_index_from_flanks(fqb::synth_flanks(s->ref), with_rollhash != 0, idx);
    if (!fqb::dump_index(idx, prefix, err)) { fqb::set_error(err); return FQB_ERR_IO; }
    if (!fqb::synth_write_index_side_files(s->ref, genome_path, dbsnp_path, prefix, err)) { fqb::set_error(err); return FQB_ERR_IO; }
    return FQB_OK;
}
int fqb_index_from_flank_fasta(const char *flank_fasta, const char *prefix, int with_rollhash) {
    std::string err;
    std::vector<fqb::FlankSeq> flanks;
    if (!fqb::read_flank_fasta(flank_fasta, flanks, err)) { fqb::set_error(err); return FQB_ERR_IO; }
    fqb::HostIndex idx;
    fqb::build_index_from_flanks(flanks, with_rollhash != 0, idx);
    if (!fqb::dump_index(idx, prefix, err)) { fqb::set_error(err); return FQB_ERR_IO; }
    return FQB_OK;
}
int fqb_synth_reads(const fqb_synth *s, const fqb_synth_read_cfg_t *c, int64_t first_pair, int64_t n_pairs,
                    uint8_t *b1, uint8_t *q1, uint8_t *b2, uint8_t *q2, int n_threads) {
    if (!s || !c || c->read_len < 35 || c->read_len > FQB_MAX_READ_LEN) { fqb::set_error("bad read config"); return FQB_ERR_ARG; }
    fqb::SynthReadConfig cfg;
    cfg.seed = c->seed; cfg.read_len = c->read_len; cfg.max_indel_len = c->max_indel_len; cfg.f_on = c->f_on;
    cfg.sub_rate = c->sub_rate; cfg.ins_rate = c->ins_rate; cfg.del_rate = c->del_rate; cfg.n_rate = c->n_rate;
    cfg.isize_mean = c->isize_mean; cfg.isize_sd = c->isize_sd; cfg.bad_tail_rate = c->bad_tail_rate;
    fqb::synth_reads(s->ref, cfg, first_pair, n_pairs, b1, q1, b2, q2, n_threads);
    return FQB_OK;
}
int fqb_write_fastq_gz(const char *path, int which_end, int64_t first_pair, int64_t n_pairs, int32_t L,
                       const uint8_t *bases, const uint8_t *quals) {
    gzFile g = gzopen(path, "wb1");
    if (!g) { fqb::set_error(std::string("cannot write ") + path); return FQB_ERR_IO; }
    gzbuffer(g, 1 << 20);
    std::string rec;
    char name[64];
    for (int64_t i = 0; i < n_pairs; ++i) {
        int n = snprintf(name, sizeof name, "@r%011lld/%d\n", (long long)(first_pair + i), which_end);
        rec.assign(name, (size_t)n);
        rec.append(reinterpret_cast<const char *>(bases + i * L), (size_t)L);
        rec += "\n+\n";
        rec.append(reinterpret_cast<const char *>(quals + i * L), (size_t)L);
        rec += "\n";
        if (gzwrite(g, rec.data(), (unsigned)rec.size()) <= 0) { gzclose(g); fqb::set_error("gzwrite failed"); return FQB_ERR_IO; }
    }
    gzclose(g);
    return FQB_OK;
}

// ---- packed input form (north_star: 2-bit bases, 128-bit loads on the device) ----
int32_t fqb_packed_stride(int32_t stride) { return stride <= 0 ? 0 : ((stride + 63) / 64) * 16; }

int fqb_pack_reads(int64_t n, int32_t stride, const uint8_t *bases, const uint8_t *quals, int32_t packed_stride, uint8_t *packed_out, uint8_t *quals_out) {
    if (n < 0 || stride < 1 || !bases || !quals || !packed_out || !quals_out || packed_stride != fqb_packed_stride(stride)) {
        fqb::set_error("fqb_pack_reads: bad arguments (packed_stride must be fqb_packed_stride(stride))"); return FQB_ERR_ARG;
    }
    // nst_nt4_table (libbwa/bntseq.c:38-55) over the bytes a FASTQ line can hold
    uint8_t nt4[256];
    for (int c = 0; c < 256; ++c) nt4[c] = 4;
    nt4['A'] = nt4['a'] = 0; nt4['C'] = nt4['c'] = 1; nt4['G'] = nt4['g'] = 2; nt4['T'] = nt4['t'] = 3; nt4['-'] = 5;
    for (int64_t r = 0; r < n; ++r) {
        const uint8_t *b = bases + (size_t)r * stride, *q = quals + (size_t)r * stride;
        uint32_t *w = reinterpret_cast<uint32_t *>(packed_out + (size_t)r * packed_stride);
        uint8_t *qo = quals_out + (size_t)r * stride;
        uint32_t q_or = 0;
        for (int k = 0, j = 0; k < packed_stride / 4; ++k) {   // one word (16 bases) at a time
            uint32_t acc = 0;
            const int end = j + 16 < stride ? j + 16 : stride;
            for (int sh = 0; j < end; ++j, sh += 2) {
                const uint32_t c = nt4[b[j]];
                acc |= (c & 3u) << sh;                        // c - 4 for the codes above 3: 0 = N, 1 = '-'
                q_or |= q[j];
                qo[j] = (uint8_t)(q[j] | ((c & 4u) << 5));    // codes 4 and 5 -> bit 7
            }
            w[k] = acc;
        }
        if (q_or & 0x80u) { fqb::set_error("fqb_pack_reads: a quality byte above 127"); return FQB_ERR_ARG; }
    }
    return FQB_OK;
}

}  // extern "C"

std::vector<fqb::FlankSeq> fqb_synth_flanks_internal(const fqb_synth *s) { return fqb::synth_flanks(s->ref); }
