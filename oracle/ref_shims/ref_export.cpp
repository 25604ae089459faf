// TEST INFRASTRUCTURE ONLY (oracle/).  Flat C exports over the reference's own
// align-stage code, for differential testing.  BwtMapper.cpp is compiled IN
// PLACE from the read-only reference tree (its hot-path functions are file
// statics), and this file replays the per-batch call sequence of
// BwtMapper::PairEndMapper (src/BwtMapper.cpp:1893-2104) single-threaded, with a
// snapshot after every stage.  No reference source is copied here.
#include FQREF_ROOT_SRC

#include "../../include/fastquick_b200.h"
#include <string>
#include <vector>

namespace {

struct Snap {                       // one snapshot of both ends
    std::vector<fqb_read_t> r[2];
};

struct Ctx {
    BwtIndexer *idx = nullptr;
    gap_opt_t *opt = nullptr;
    pe_opt_t *popt = nullptr;
    bwa_seqio_t *ks[2] = {nullptr, nullptr};
    bwa_seq_t *seqs[2] = {nullptr, nullptr};
    int cap = 0;
    int n = 0;
    uint32_t round = 0;
    isize_info_t last_ii, ii;
    kh_64_t *hash = nullptr;
    ubyte_t *pacseq = nullptr;
    Snap prep, pe, sw, fin;
    std::vector<int32_t> aln_off[2];
    std::vector<fqb_aln_t> aln[2];
    std::vector<uint8_t> seq_codes[2];   // forward-orientation nt4 codes after prep, cap*read_len
    std::string md[2];                   // '\n' separated per read
    std::vector<uint32_t> multi[2];      // per read FQB_MAX_MULTI * 2 words: pos, (gap<<16|mm<<8|strand)
    StatCollector *collector = nullptr;
    std::ofstream *fout = nullptr;
    FileStatCollector *fsc = nullptr;
    std::string out_prefix;
};

void snap(Ctx *c, Snap &s) {
    for (int e = 0; e < 2; ++e) {
        s.r[e].assign(c->n, fqb_read_t());
        for (int i = 0; i < c->n; ++i) {
            const bwa_seq_t *p = c->seqs[e] + i;
            fqb_read_t &o = s.r[e][i];
            memset(&o, 0, sizeof(o));
            o.pos = p->pos; o.sa = p->sa; o.c1 = p->c1; o.c2 = p->c2; o.score = p->score;
            o.len = p->len; o.full_len = p->full_len; o.clip_len = p->clip_len;
            o.type = p->type; o.strand = p->strand; o.filtered = p->filtered; o.extra_flag = p->extra_flag;
            o.n_mm = p->n_mm; o.n_gapo = p->n_gapo; o.n_gape = p->n_gape; o.mapQ = p->mapQ;
            o.seQ = p->seQ; o.n_multi = p->n_multi; o.nm = p->nm; o.n_aln = p->n_aln;
            o.has_cigar = p->cigar != 0;
            o.n_cigar = p->cigar ? p->n_cigar : 0;
            for (int k = 0; k < o.n_cigar && k < FQB_MAX_CIGAR; ++k) o.cigar[k] = p->cigar[k];
        }
    }
}

}  // namespace

extern "C" {

void *fqref_open(const char *index_prefix, int kmer_thresh, int trim_qual, int batch_cap) {
    Ctx *c = new Ctx();
    c->opt = gap_init_opt();
    c->popt = bwa_init_pe_opt();
    c->opt->trim_qual = trim_qual;
    c->idx = new BwtIndexer(kmer_thresh);
    std::string prefix(index_prefix);
    // the subset of runAlign's .param parsing that the aligner itself needs (src/FASTQuick.cpp:365-467)
    {
        std::ifstream par(prefix + ".param");
        std::string k, v;
        while (par >> k >> v) {
            if (k == "NUM_VAR_LONG") c->opt->num_variant_long = atoi(v.c_str());
            else if (k == "NUM_VAR_SHORT") c->opt->num_variant_short = atoi(v.c_str());
            else if (k == "SHORT_FLANK_LENGTH") c->opt->flank_len = atoi(v.c_str());
            else if (k == "LONG_FLANK_LENGTH") c->opt->flank_long_len = atoi(v.c_str());
        }
    }
    c->idx->LoadIndex(prefix);
    c->cap = batch_cap > 0 ? batch_cap : READ_BUFFER_SIZE;
    for (int e = 0; e < 2; ++e) {
        c->seqs[e] = (bwa_seq_t *)calloc(c->cap, sizeof(bwa_seq_t));
        bwa_init_read_seq(c->cap, c->seqs[e], c->opt);
    }
    c->last_ii.avg = -1.0;
    c->hash = kh_init(64);
    bwase_initialize();
    srand48(c->idx->bns->seed);
    return c;
}

gap_opt_t *fqref_gap_opt(void *h) { return ((Ctx *)h)->opt; }
pe_opt_t *fqref_pe_opt(void *h) { return ((Ctx *)h)->popt; }

// StatCollector set-up as in BwtMapper::BwtMapper (src/BwtMapper.cpp:225-230)
int fqref_open_stats(void *h, const char *index_prefix, const char *out_prefix) {
    Ctx *c = (Ctx *)h;
    c->collector = new StatCollector();
    c->collector->RestoreVcfSites(index_prefix, c->opt);
    c->collector->SetGenomeSize(c->idx->ref_genome_size, c->idx->ref_N_size);
    c->out_prefix = out_prefix;
    c->fout = new std::ofstream(c->out_prefix + ".InsertSizeTable");
    c->fsc = new FileStatCollector("r1.fq", "r2.fq");
    return 0;
}

int fqref_open_reads(void *h, const char *fq1, const char *fq2) {
    Ctx *c = (Ctx *)h;
    c->ks[0] = bwa_seq_open(fq1);
    c->ks[1] = bwa_seq_open(fq2);
    c->round = 0;
    srand48(c->idx->bns->seed);            // PairEndMapper re-seeds per file pair (src/BwtMapper.cpp:1817)
    c->last_ii.avg = -1.0;
    return (c->ks[0] && c->ks[1]) ? 0 : -1;
}

// One batch through every stage.  Returns the number of pairs (0 at EOF).
int fqref_next_batch(void *h) {
    Ctx *c = (Ctx *)h;
    bwt_t *bwt[2] = {c->idx->bwt_d, c->idx->rbwt_d};
    int n0 = 0, n1 = 0;
    for (int e = 0; e < 2; ++e) bwa_clean_read_seq(c->n, c->seqs[e]);
    int r0 = bwa_read_seq_with_hash_dev(c->idx, c->ks[0], c->cap, &n0, c->opt->mode, c->opt->trim_qual, c->opt->frac,
                                        c->round, c->seqs[0], c->opt->read_len);
    int r1 = bwa_read_seq_with_hash_dev(c->idx, c->ks[1], c->cap, &n1, c->opt->mode, c->opt->trim_qual, c->opt->frac,
                                        c->round, c->seqs[1], c->opt->read_len);
    c->round++;
    if (r0 == 0 || r1 == 0 || n0 != n1) { c->n = 0; return 0; }
    c->n = n0;
    snap(c, c->prep);
    const int L = c->opt->read_len;
    for (int e = 0; e < 2; ++e) {          // forward-orientation codes: filtered reads were never reversed
        c->seq_codes[e].assign((size_t)c->n * L, 4);
        for (int i = 0; i < c->n; ++i) {
            const bwa_seq_t *p = c->seqs[e] + i;
            for (int j = 0; j < (int)p->len; ++j)
                c->seq_codes[e][(size_t)i * L + j] = p->filtered ? p->seq[j] : p->seq[p->len - 1 - j];
        }
    }
    // (1) bwa_cal_sa_reg_gap over each end (src/BwtMapper.cpp:1933-1952)
    for (int e = 0; e < 2; ++e) {
        bwa_cal_sa_reg_gap(0, bwt, c->n, c->seqs[e], c->opt, c->idx);
        c->aln_off[e].assign(c->n + 1, 0);
        c->aln[e].clear();
        for (int i = 0; i < c->n; ++i) {
            const bwa_seq_t *p = c->seqs[e] + i;
            for (int k = 0; k < p->n_aln; ++k) {
                fqb_aln_t a;
                a.k = p->aln[k].k; a.l = p->aln[k].l; a.score = p->aln[k].score;
                a.n_mm = p->aln[k].n_mm; a.n_gapo = p->aln[k].n_gapo; a.n_gape = p->aln[k].n_gape; a.a = p->aln[k].a;
                c->aln[e].push_back(a);
            }
            c->aln_off[e][i + 1] = (int32_t)c->aln[e].size();
        }
    }
    // (2) PEworker body (src/BwtMapper.cpp:654-684)
    bwa_cal_pac_pos_pe(bwt, c->n, c->seqs, &c->ii, c->popt, c->opt, &c->last_ii, c->hash);
    snap(c, c->pe);
    for (int e = 0; e < 2; ++e) {
        c->multi[e].assign((size_t)c->n * FQB_MAX_MULTI * 2, 0);
        for (int i = 0; i < c->n; ++i) {
            const bwa_seq_t *p = c->seqs[e] + i;
            for (int k = 0; k < p->n_multi && k < FQB_MAX_MULTI; ++k) {
                c->multi[e][((size_t)i * FQB_MAX_MULTI + k) * 2] = p->multi[k].pos;
                c->multi[e][((size_t)i * FQB_MAX_MULTI + k) * 2 + 1] =
                    (uint32_t)p->multi[k].gap << 16 | (uint32_t)p->multi[k].mm << 8 | p->multi[k].strand;
            }
        }
    }
    c->pacseq = bwa_paired_sw(c->idx->bns, c->pacseq ? c->pacseq : c->idx->pac_buf, c->n, c->seqs, c->popt, &c->ii, c->opt->mode);
    snap(c, c->sw);
    for (int e = 0; e < 2; ++e) bwa_refine_gapped(c->idx->bns, c->n, c->seqs[e], c->pacseq, 0);
    snap(c, c->fin);
    for (int e = 0; e < 2; ++e) {
        c->md[e].clear();
        for (int i = 0; i < c->n; ++i) {
            const bwa_seq_t *p = c->seqs[e] + i;
            if (p->md) c->md[e] += p->md;
            c->md[e] += '\n';
        }
    }
    c->last_ii = c->ii;
    // (3) statistics, main-thread loop of PairEndMapper (src/BwtMapper.cpp:2053-2085), BAM emission omitted
    if (c->collector) {
        FileStatCollector &FSC = *c->fsc;
        for (int i = 0; i < c->n; ++i) {
            bwa_seq_t *p[2] = {c->seqs[0] + i, c->seqs[1] + i};
            FSC.NumBase += p[0]->full_len;
            FSC.NumBase += p[1]->full_len;
            if (p[0]->filtered && p[1]->filtered) { ++FSC.TotalFiltered; continue; }
            if (p[0]->type == BWA_TYPE_NO_MATCH && p[1]->type == BWA_TYPE_NO_MATCH) { ++FSC.BwaUnmapped; continue; }
            FSC.TotalRetained += c->collector->AddAlignment(c->idx->bns, p[0], p[1], c->opt, *c->fout, FSC.TotalMAPQ);
        }
        FSC.NumRead += 2 * (long long)c->n;
    }
    return c->n;
}

// stage: 0 prep, 1 after bwa_cal_pac_pos_pe, 2 after bwa_paired_sw, 3 after bwa_refine_gapped
const fqb_read_t *fqref_rows(void *h, int stage, int end) {
    Ctx *c = (Ctx *)h;
    Snap *s = stage == 0 ? &c->prep : stage == 1 ? &c->pe : stage == 2 ? &c->sw : &c->fin;
    return s->r[end].data();
}
const int32_t *fqref_aln_off(void *h, int end) { return ((Ctx *)h)->aln_off[end].data(); }
const fqb_aln_t *fqref_aln(void *h, int end) { return ((Ctx *)h)->aln[end].data(); }
const uint8_t *fqref_seq_codes(void *h, int end) { return ((Ctx *)h)->seq_codes[end].data(); }
const char *fqref_md(void *h, int end) { return ((Ctx *)h)->md[end].c_str(); }
const uint32_t *fqref_multi(void *h, int end) { return ((Ctx *)h)->multi[end].data(); }
void fqref_isize(void *h, fqb_isize_t *o) {
    Ctx *c = (Ctx *)h;
    o->avg = c->ii.avg; o->std = c->ii.std; o->ap_prior = c->ii.ap_prior;
    o->low = c->ii.low; o->high = c->ii.high; o->high_bayesian = c->ii.high_bayesian; o->pad_ = 0;
}
int fqref_read_len(void *h) { return ((Ctx *)h)->opt->read_len; }

// FileStatCollector counters: NumRead NumBase TotalFiltered BwaUnmapped TotalMAPQ TotalRetained
void fqref_fsc(void *h, long long *out6) {
    FileStatCollector &F = *((Ctx *)h)->fsc;
    out6[0] = F.NumRead; out6[1] = F.NumBase; out6[2] = F.TotalFiltered; out6[3] = F.BwaUnmapped; out6[4] = F.TotalMAPQ; out6[5] = F.TotalRetained;
}
int fqref_finish_stats(void *h) {
    Ctx *c = (Ctx *)h;
    if (!c->collector) return -1;
    c->fout->close();
    c->collector->AddFSC(*c->fsc);
    c->collector->ProcessCore(c->out_prefix, c->opt);
    return 0;
}

// direct taps on single functions
int fqref_maxdiff(int l, double err, double thres) { return bwa_cal_maxdiff(l, err, thres); }
uint32_t fqref_bwt_sa(void *h, int which, uint32_t k) { Ctx *c = (Ctx *)h; return bwt_sa(which ? c->idx->rbwt_d : c->idx->bwt_d, k); }
int fqref_is_filtered(void *h, const uint8_t *codes, int len) { return ((Ctx *)h)->idx->IsReadFiltered((ubyte_t *)codes, 0, len) ? 1 : 0; }

}  // extern "C"
