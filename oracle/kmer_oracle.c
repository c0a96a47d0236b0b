/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the PhaMers k-mer counting loop, used by the
 * parity tests at sizes where the Python restatement (oracle/phamers_oracle.py) is too slow.
 * Never linked into, loaded by, or called from the product library (phamers_b200/).
 *
 * Parity status: PINNED -- tests/test_oracle.py checks it bin-for-bin against phamers_oracle.py
 * (itself pinned to the unmodified reference and to tests/golden/).
 *
 * Restates /root/reference/scripts/kmer.py:
 *   :183-196  sequence_to_integers -- only the bytes 'A','T','G','C' (upper case) are symbols, with
 *             indices A=0 T=1 G=2 C=3; any other byte is the blank '-'.
 *   :42-50    count_string -- for every start i in [0, L-k], the window s[i:i+k] is counted in bin
 *             int(window, 4) (first base most significant) unless it holds a blank.
 *   :209-221  normalize_counts -- row / row-sum in float64 (0/0 = NaN for an empty row).
 */
#include <stdint.h>
#include <string.h>
#include <math.h>

static inline int sym(uint8_t c) {
    switch (c) { case 'A': return 0; case 'T': return 1; case 'G': return 2; case 'C': return 3; default: return -1; }
}

/* counts[n][4^k] (int64) for n sequences laid end to end in `seq`, sequence i = seq[off[i] .. off[i+1]). */
int phm_oracle_count(const uint8_t *seq, const int64_t *off, int64_t n, int k, int64_t *counts) {
    if (k < 1 || k > 12) return -1;
    const int64_t bins = (int64_t)1 << (2 * k);
    const uint64_t mask = (uint64_t)bins - 1;
    memset(counts, 0, (size_t)(n * bins) * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) {
        int64_t *row = counts + i * bins;
        uint64_t idx = 0;
        int run = 0;                       /* consecutive symbols ending at the current base */
        for (int64_t p = off[i]; p < off[i + 1]; ++p) {
            int s = sym(seq[p]);
            if (s < 0) { run = 0; idx = 0; continue; }
            idx = ((idx << 2) | (uint64_t)s) & mask;
            if (++run >= k) row[idx] += 1;
        }
    }
    return 0;
}

/* freq[n][bins] = counts / row sum, float64 (kmer.py:209-221; NaN for an all-zero row). */
int phm_oracle_normalize(const int64_t *counts, int64_t n, int64_t bins, double *freq) {
    for (int64_t i = 0; i < n; ++i) {
        int64_t total = 0;
        for (int64_t b = 0; b < bins; ++b) total += counts[i * bins + b];
        for (int64_t b = 0; b < bins; ++b)
            freq[i * bins + b] = total ? (double)counts[i * bins + b] / (double)total : NAN;
    }
    return 0;
}
