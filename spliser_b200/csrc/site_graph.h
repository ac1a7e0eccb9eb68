// Host-side site table + competing-site graph (structure only; the counts are reduced on the GPU).
// Replaces the Python object graph of Site instances (Gene_Site_Iter_Graph_v0_1_8.py:98-339) built
// by findAlphaCounts (SpliSER_v0_1_8.py:227-362) and findCompetitorPos (:364-372) with flat
// arrays + CSR, in the reference's own list order so that results need no re-ordering.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace spl {

// strand class of a site as seen by check_strand (SpliSER_v0_1_8.py:374-406)
enum : uint8_t {
    CLS_ANY = 0,     // unstranded run: every read matches
    CLS_PLUS = 1,    // stranded run, site strand '+'
    CLS_MINUS = 2,   // stranded run, site strand '-'
    CLS_NEVER = 3,   // stranded run, any other strand byte ('?', '' ...): no read ever matches
    CLS_PSEUDO = 4   // recount only: a partner position that is not itself a gap site
};

struct SiteGraph {
    int32_t n_chrom = 0;
    int64_t n_sites = 0;
    std::vector<int64_t> cs_off;       // [n_chrom+1] site range of each chromosome
    std::vector<int32_t> chrom;        // [S]
    std::vector<int32_t> pos;          // [S] non-decreasing inside a chromosome
    std::vector<uint8_t> strand;       // [S] raw byte shown in the TSV (first-seen strand)
    std::vector<uint8_t> cls;          // [S] CLS_*
    std::vector<int64_t> first_line;   // [S] junction-table row that created the site

    // Site.Partners (objects, first-appearance order, G:260-262)
    std::vector<int64_t> pt_off;       // [S+1]
    std::vector<int32_t> pt_site;      // partner site index
    // Site.PartnerCounts keys (positions, insertion order, G:243-246); counts live on the device
    std::vector<int64_t> pc_off;       // [S+1]
    std::vector<int32_t> pc_pos;
    // Site.CompetitorPos (sorted unique positions, G:264-266)
    std::vector<int64_t> cp_off;       // [S+1]
    std::vector<int32_t> cp_pos;
    // reverse partner index: rp lists hang off the FIRST site index of each distinct position e
    // and hold every site t of that chromosome with e in PartnerCounts(t)
    std::vector<int64_t> rp_off;       // [S+1]
    std::vector<int32_t> rp_site;

    // segments for the alpha / partner-count reduction (K1): junction rows grouped by site / by
    // PartnerCounts entry
    std::vector<int64_t> inc_off;      // [S+1]
    std::vector<int32_t> inc_line;     // rows whose score adds to alpha of the site (S:341)
    std::vector<int64_t> einc_off;     // [E+1], E = pc_pos.size()
    std::vector<int32_t> einc_line;    // rows whose score adds to that PartnerCounts entry (S:353-355)

    // recount only: map from table row to caller's gap index (-1 for pseudo sites)
    std::vector<int64_t> gap_index;
    bool dirty_regime = false;
};

// process: junction table in BED-line order (after the text filters).  Returns "" or an error.
std::string build_site_graph(int32_t n_chrom, int64_t n_junc, const int32_t* j_chrom, const int32_t* j_left,
                             const int32_t* j_right, const uint8_t* j_strand, bool stranded, SiteGraph& g);

// combine re-count: gap sites with explicit partner / competitor position lists.
std::string build_recount_graph(int32_t n_chrom, int64_t n_sites, const int32_t* s_chrom, const int32_t* s_pos,
                                const uint8_t* s_strand, const int64_t* p_off, const int32_t* p_pos,
                                const int64_t* c_off, const int32_t* c_pos, bool stranded, SiteGraph& g);

}  // namespace spl
