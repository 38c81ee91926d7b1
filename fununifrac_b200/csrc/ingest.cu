// ingest.cu — native ingest of sourmash `gather` CSV profiles (host code; no device work).
//
// Replaces the per-file work of the reference's ingest pool — pd.read_csv of all 29 columns followed by a Python
// iterrows() loop (functional_profile_to_vector, src/factory/make_emd_input.py:82-122; pool :58-79) — with one pass
// over the file bytes that tokenises only the two columns the path uses (`name` and the abundance key), looks the KO
// up in a hash of the tree's node names and scatters into the dense profile row.  A thread pool runs over the files.
//
// Parity contract (pinned in tests/test_ingest_native.py against pandas + the Python mirror of the reference):
//   * fields: RFC-4180 quoting as pandas' C tokenizer does it (delimiter ',', quote '"', doubled quotes, \r\n or \n,
//     blank lines skipped, first non-blank line = header, short rows padded with NA, long rows = error);
//   * numbers: pandas' default float parser for the C engine (`precise_xstrtod`: up to 17 significant digits
//     accumulated in a double, one multiplication or division by an exact power of ten), its NA strings -> NaN;
//     anything else in the abundance column is "not a number" (the reference raises ValueError, :99-103);
//   * KO id: name.split(':')[-1] if the name starts with 'ko:' else name.split('|')[-1].split(':')[-1] (:105-109);
//     unknown ids are reported back (the reference prints a warning and skips, :110-111); duplicate ids: the last
//     row wins (plain assignment, :113);
//   * P.sum() == 0 -> "empty profile" error (:117-119); P / P.sum() with numpy's pairwise summation (:120-121);
//     --unweighted: P[P > 0] = 1 after the normalisation (:62-64).
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <deque>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include <immintrin.h>

#include "common.cuh"

struct fuf_name_index {
    std::deque<std::string> names;                            // owns the text the views of `map` point into
    std::unordered_map<std::string_view, int32_t> map;        // looked up with views into the file buffer: no copies
    int32_t n = 0;
    // Names of at most 8 bytes (KO ids are 6: "K01234") also live in a flat open-addressing table keyed by the bytes
    // themselves packed into one 64-bit word: one multiply, one cache line, no string comparison.  (key 0 = empty slot; a
    // packed name is never 0 because names are not empty and hold no NUL bytes.)
    std::vector<uint64_t> flat_key;
    std::vector<int32_t> flat_val;
    uint64_t flat_mask = 0;
    static inline uint64_t pack(const char *p, size_t len)
    {
        uint64_t w = 0;
        memcpy(&w, p, len);            // len <= 8
        return w | ((uint64_t)len << 60);   // len <= 7: the top byte is free for the length
    }
    static inline uint64_t mix(uint64_t w) { return (w * 0x9E3779B97F4A7C15ull) >> 20; }
    // index of the name, or -1
    inline int32_t find(const char *p, size_t len) const
    {
        if (len <= 7 && len > 0 && flat_mask) {
            const uint64_t key = pack(p, len);
            for (uint64_t h = mix(key) & flat_mask;; h = (h + 1) & flat_mask) {
                if (flat_key[h] == key) return flat_val[h];
                if (flat_key[h] == 0) break;
            }
            return -1;
        }
        auto it = map.find(std::string_view(p, len));
        return it == map.end() ? -1 : it->second;
    }
};

namespace fuf {
namespace ingest {

// numpy's float64 add-reduce over a contiguous vector: pairwise summation, 8-way unrolled 128-element base case.
static double np_pairwise_sum(const double *a, ptrdiff_t n)
{
    if (n < 8) {
        double res = 0.;
        for (ptrdiff_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        ptrdiff_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int u = 0; u < 8; u++) r[u] += a[i + u];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    ptrdiff_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

static const double *pow10_table()
{
    static double e[309];
    static bool init = false;
    if (!init) {
        char buf[16];
        for (int i = 0; i <= 308; i++) {
            snprintf(buf, sizeof(buf), "1e%d", i);
            e[i] = strtod(buf, nullptr);  // correctly rounded literals, as the compiler makes of pandas' table
        }
        init = true;
    }
    return e;
}

// pandas/_libs/src/parser/tokenizer.c: precise_xstrtod with decimal '.', sci 'E', no thousands separator, trailing
// whitespace skipped.  Returns false if [s, s+len) is not entirely one number.
static bool parse_double(const char *s, size_t len, double *out)
{
    const double *e = pow10_table();
    const char *p = s, *end = s + len;
    auto digit = [&](const char *q) { return q < end && *q >= '0' && *q <= '9'; };
    while (p < end && (*p == ' ' || (*p >= '\t' && *p <= '\r'))) p++;
    bool negative = false;
    if (p < end && (*p == '-' || *p == '+')) negative = (*p++ == '-');
    double number = 0.;
    int exponent = 0, num_digits = 0, num_decimals = 0;
    const int max_digits = 17;
    while (digit(p)) {
        if (num_digits < max_digits) {
            number = number * 10. + (*p - '0');
            num_digits++;
        } else {
            ++exponent;
        }
        p++;
    }
    if (p < end && *p == '.') {
        p++;
        while (num_digits < max_digits && digit(p)) {
            number = number * 10. + (*p - '0');
            p++;
            num_digits++;
            num_decimals++;
        }
        if (num_digits >= max_digits)
            while (digit(p)) ++p;
        exponent -= num_decimals;
    }
    if (num_digits == 0) return false;
    if (negative) number = -number;
    if (p < end && (*p == 'E' || *p == 'e')) {
        const char *q = p + 1;
        bool eneg = false;
        if (q < end && (*q == '-' || *q == '+')) eneg = (*q++ == '-');
        int n = 0, nd = 0;
        while (digit(q) && nd < max_digits) {
            n = n * 10 + (*q - '0');
            nd++;
            q++;
        }
        if (nd > 0) {  // pandas leaves p on the 'E' (and then fails the whole-field check) when no digits follow
            exponent += eneg ? -n : n;
            p = q;
        }
    }
    if (exponent > 308) return false;  // pandas: ERANGE -> the column is not numeric
    if (exponent > 0) {
        number *= e[exponent];
    } else if (exponent < -308) {
        if (exponent < -616) {
            number = 0.;
        } else {
            number /= e[-308 - exponent];
            number /= e[308];
        }
    } else {
        number /= e[-exponent];
    }
    if (number == HUGE_VAL || number == -HUGE_VAL) return false;
    while (p < end && (*p == ' ' || (*p >= '\t' && *p <= '\r'))) p++;
    if (p != end) return false;
    *out = number;
    return true;
}

static bool is_na(const char *s, size_t len)
{
    static const char *na[] = {"", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN",
                               "<NA>", "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"};
    if (len == 0) return true;
    // every NA spelling is at most 8 characters long and starts with one of  # - 1 < N n : nearly every real field is
    // rejected here without a string comparison
    if (len > 8) return false;
    const char c = s[0];
    if (c != '#' && c != '-' && c != '1' && c != '<' && c != 'N' && c != 'n') return false;
    if (c == '1' && !(len >= 6 && s[1] == '.' && s[2] == '#')) return false;
    for (const char *w : na)
        if (strlen(w) == len && memcmp(w, s, len) == 0) return true;
    return false;
}

// Non-NA text pandas' float conversion accepts besides plain numbers: infinities, any case.
static bool parse_special(const char *s, size_t len, double *out)
{
    std::string t(s, len);
    bool neg = false;
    size_t i = 0;
    if (i < t.size() && (t[i] == '-' || t[i] == '+')) neg = (t[i++] == '-');
    std::string r = t.substr(i);
    for (auto &c : r) c = (char)tolower((unsigned char)c);
    if (r == "inf" || r == "infinity") {
        *out = neg ? -INFINITY : INFINITY;
        return true;
    }
    return false;
}

struct Field {
    const char *p;
    size_t len;
    bool quoted;
};

// Set when a record ended on a '\r' that no '\n' follows (classic Mac line ends).  pandas' C tokenizer has quirks of its
// own there (with quoted fields it can fail with "Buffer overflow caught" or shift columns), so such a file is handed to
// the pandas route as a whole: whatever it returns or raises is then the reference's behaviour by construction.
static thread_local bool t_lone_cr = false;

// First ',', '\n' or '\r' at or after p (end if none): eight bytes at a time.
static inline const char *skip_plain(const char *p, const char *end)
{
    const uint64_t ones = 0x0101010101010101ull, high = 0x8080808080808080ull;
    while (p + 8 <= end) {
        uint64_t w;
        memcpy(&w, p, 8);
        const uint64_t a = w ^ (ones * (uint64_t)','), b = w ^ (ones * (uint64_t)'\n'), c = w ^ (ones * (uint64_t)'\r');
        const uint64_t m = (((a - ones) & ~a) | ((b - ones) & ~b) | ((c - ones) & ~c)) & high;
        if (m) return p + (__builtin_ctzll(m) >> 3);   // lowest flagged byte = first delimiter (little endian)
        p += 8;
    }
    while (p < end && *p != ',' && *p != '\n' && *p != '\r') p++;
    return p;
}

// One record starting at *cur, tokenised with pandas' C-parser rules; `sink(index, field)` receives every field whose
// text `want(index)` asks for (the others are only counted).  Unescaped text of a wanted quoted field with doubled
// quotes lives in `scratch`.  Returns false at end of input; *n_fields = fields in the record, *first_empty = the first
// field is empty and unquoted (a record with one such field is a blank line).
template <class Want, class Sink>
static bool tokenize_record(const char *&cur, const char *end, std::deque<std::string> &scratch, size_t *n_fields,
                            bool *first_empty, Want want, Sink sink)
{
    scratch.clear();
    *n_fields = 0;
    *first_empty = false;
    if (cur >= end) return false;
    auto emit = [&](Field f) {
        const size_t idx = (*n_fields)++;
        if (idx == 0) *first_empty = f.len == 0 && !f.quoted;
        if (want(idx)) sink(idx, f);
    };
    for (;;) {
        Field f{cur, 0, false};
        const bool keep = want(*n_fields);
        if (cur < end && *cur == '"') {
            f.quoted = true;
            const char *q = cur + 1;
            bool escaped = false;
            const char *start = q;
            while (q < end) {
                if (*q == '"') {
                    if (q + 1 < end && q[1] == '"') {
                        escaped = true;
                        q += 2;
                        continue;
                    }
                    break;
                }
                q++;
            }
            const char *close = q;
            f.p = start;
            f.len = (size_t)(close - start);
            if (escaped && keep) {
                std::string s;
                for (const char *r = start; r < close; r++) {
                    s.push_back(*r);
                    if (*r == '"') r++;
                }
                scratch.push_back(std::move(s));
                f.p = nullptr;  // resolved below
                f.len = scratch.size() - 1;
            }
            cur = close < end ? close + 1 : end;
            // pandas appends whatever follows the closing quote up to the delimiter
            const char *tail = cur;
            cur = skip_plain(cur, end);
            if (cur > tail && keep) {
                std::string s = (f.p ? std::string(f.p, f.len) : scratch[f.len]) + std::string(tail, cur - tail);
                scratch.push_back(std::move(s));
                f.p = nullptr;
                f.len = scratch.size() - 1;
            }
            if (keep && !f.p) {
                const std::string &s = scratch[f.len];
                f.p = s.data();
                f.len = s.size();
            }
        } else {
            cur = skip_plain(cur, end);
            f.len = (size_t)(cur - f.p);
        }
        emit(f);
        if (cur >= end) break;
        if (*cur == ',') {
            cur++;
            if (cur >= end) {  // trailing delimiter: one more empty field
                emit(Field{cur, 0, false});
                break;
            }
            continue;
        }
        if (*cur == '\r') {
            cur++;
            if (cur < end && *cur == '\n') cur++;
            else t_lone_cr = true;
        } else {
            cur++;  // '\n'
        }
        break;
    }
    return true;
}

// Fast path for the data rows of a gather CSV: a record without any quote character is just its text up to the first
// '\n' / '\r' split at the commas.  32 bytes at a time (AVX2, checked at run time), the offsets of all commas are
// collected; the two wanted fields are then read off directly.  Returns false — nothing consumed — when the record
// holds a quote (pandas' quoting rules live in tokenize_record), has more commas than the buffer, or the CPU has no AVX2.
constexpr int FAST_MAX_COMMAS = 256;
struct FastRow {
    const char *row_end, *next;
    int n_commas;
    uint32_t comma[FAST_MAX_COMMAS];
};
__attribute__((target("avx2"))) static bool scan_row_avx2(const char *cur, const char *end, FastRow &r)
{
    const __m256i vc = _mm256_set1_epi8(','), vn = _mm256_set1_epi8('\n'), vr = _mm256_set1_epi8('\r'), vq = _mm256_set1_epi8('"');
    const char *p = cur;
    int nc = 0;
    const char *row_end = nullptr;
    while (p + 32 <= end) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
        uint32_t mc = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, vc));
        const uint32_t ml = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_cmpeq_epi8(v, vn), _mm256_cmpeq_epi8(v, vr)));
        uint32_t mq = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, vq));
        if (ml) {
            const uint32_t first = (uint32_t)__builtin_ctz(ml);
            const uint32_t keep = first ? (0xFFFFFFFFu >> (32 - first)) : 0u;
            mc &= keep;
            mq &= keep;
            row_end = p + first;
        }
        if (mq) return false;
        if (nc + __builtin_popcount(mc) > FAST_MAX_COMMAS) return false;
        const uint32_t base = (uint32_t)(p - cur);
        while (mc) {
            r.comma[nc++] = base + (uint32_t)__builtin_ctz(mc);
            mc &= mc - 1;
        }
        if (row_end) break;
        p += 32;
    }
    if (!row_end) {
        for (; p < end; p++) {
            const char c = *p;
            if (c == '\n' || c == '\r') break;
            if (c == '"') return false;
            if (c == ',') {
                if (nc >= FAST_MAX_COMMAS) return false;
                r.comma[nc++] = (uint32_t)(p - cur);
            }
        }
        row_end = p;
    }
    r.row_end = row_end;
    r.n_commas = nc;
    const char *next = row_end;
    if (next < end) {
        if (*next == '\r') {
            next++;
            if (next < end && *next == '\n') next++;
            else t_lone_cr = true;
        } else {
            next++;
        }
    }
    r.next = next;
    return true;
}
static inline bool cpu_has_avx2()
{
    static const bool has = __builtin_cpu_supports("avx2") && !getenv("FUF_INGEST_NO_AVX2");
    return has;
}

// All fields of a record (the header).
static bool next_record(const char *&cur, const char *end, std::vector<Field> &fields, std::deque<std::string> &scratch,
                        bool *blank)
{
    fields.clear();
    size_t n = 0;
    bool first_empty = false;
    const bool ok = tokenize_record(cur, end, scratch, &n, &first_empty, [](size_t) { return true; },
                                    [&](size_t, const Field &f) { fields.push_back(f); });
    *blank = ok && n == 1 && first_empty;
    return ok;
}

struct FileResult {
    int rc = 0;  // 0 ok; 1 not a number; 2 empty profile; 3 io / malformed; 4 bad name
    std::string message, warnings;
};

static void set_msg(FileResult &r, int rc, const std::string &m)
{
    r.rc = rc;
    r.message = m;
}

static void parse_file(const fuf_name_index *index, const char *path, const std::string &key, bool normalize, bool unweighted,
                       double *P, FileResult &res)
{
    const int32_t n = index->n;
    for (int32_t k = 0; k < n; k++) P[k] = 0.0;
    FILE *fp = fopen(path, "rb");
    if (!fp) return set_msg(res, 3, std::string("cannot open ") + path + ": " + strerror(errno));
    std::string buf;
    {   // one read of the whole file when its size is known, chunks otherwise (pipes)
        long size = -1;
        if (fseek(fp, 0, SEEK_END) == 0) {
            size = ftell(fp);
            rewind(fp);
        }
        if (size > 0) {
            buf.resize((size_t)size);
            buf.resize(fread(&buf[0], 1, (size_t)size, fp));
        }
        char chunk[1 << 16];
        size_t got;
        while ((got = fread(chunk, 1, sizeof(chunk), fp)) > 0) buf.append(chunk, got);
        fclose(fp);
    }
    const char *cur = buf.data(), *end = cur + buf.size();
    if (buf.size() >= 3 && (unsigned char)cur[0] == 0xEF && (unsigned char)cur[1] == 0xBB && (unsigned char)cur[2] == 0xBF) cur += 3;
    std::vector<Field> fields;
    std::deque<std::string> scratch;   // stable addresses: fields point into its strings
    bool blank = false;
    t_lone_cr = false;
    // header
    do {
        if (!next_record(cur, end, fields, scratch, &blank)) return set_msg(res, 3, std::string("No columns to parse from file ") + path);
    } while (blank);
    int name_col = -1, key_col = -1;
    const size_t n_cols = fields.size();
    for (size_t c = 0; c < n_cols; c++) {
        std::string h(fields[c].p, fields[c].len);
        if (h == "name" && name_col < 0) name_col = (int)c;
        if (h == key && key_col < 0) key_col = (int)c;
    }
    if (key_col < 0) return set_msg(res, 5, key);          // KeyError(abundance_key) in the reference
    if (name_col < 0) return set_msg(res, 5, "name");
    long long line = 1;
    const size_t c_name = (size_t)name_col, c_key = (size_t)key_col;
    const bool fast = cpu_has_avx2();
    FastRow row;
    for (;;) {
        // data rows: only the two columns the path uses are materialised, the rest is skipped and counted
        Field f_name{nullptr, 0, false}, f_key{nullptr, 0, false};
        bool have_name = false, have_key = false, first_empty = false;
        size_t n_fields = 0;
        if (cur >= end) break;
        if (fast && scan_row_avx2(cur, end, row)) {
            n_fields = (size_t)row.n_commas + 1;
            auto field = [&](size_t i) {
                const char *b = i == 0 ? cur : cur + row.comma[i - 1] + 1;
                const char *e = i < (size_t)row.n_commas ? cur + row.comma[i] : row.row_end;
                return Field{b, (size_t)(e - b), false};
            };
            first_empty = row.n_commas == 0 ? row.row_end == cur : row.comma[0] == 0;
            if (c_name < n_fields) f_name = field(c_name), have_name = true;
            if (c_key < n_fields) f_key = field(c_key), have_key = true;
            cur = row.next;
        } else if (!tokenize_record(cur, end, scratch, &n_fields, &first_empty,
                                    [&](size_t i) { return i == c_name || i == c_key; },
                                    [&](size_t i, const Field &f) {
                                        if (i == c_name) f_name = f, have_name = true;
                                        if (i == c_key) f_key = f, have_key = true;
                                    }))
            break;
        line++;
        if (t_lone_cr) return set_msg(res, 3, "carriage-return line ends: left to the pandas route");
        if (n_fields == 1 && first_empty) continue;        // blank line
        if (n_fields > n_cols)
            return set_msg(res, 3, "Error tokenizing data. C error: Expected " + std::to_string(n_cols) + " fields in line " +
                                       std::to_string(line) + ", saw " + std::to_string(n_fields));
        // abundance
        double value = NAN;
        if (have_key) {
            const Field &f = f_key;
            if (!(is_na(f.p, f.len) && !(f.quoted && f.len == 0 && false))) {
                if (!parse_double(f.p, f.len, &value) && !parse_special(f.p, f.len, &value))
                    return set_msg(res, 1, std::string(f.p, f.len));
            }
        }
        // name -> KO id
        if (!have_name || is_na(f_name.p, f_name.len))
            return set_msg(res, 4, "nan");
        const Field &nf = f_name;
        const char *b = nf.p, *e = nf.p + nf.len;
        if (!(nf.len >= 3 && memcmp(b, "ko:", 3) == 0)) {
            for (const char *q = e; q > b; q--)
                if (q[-1] == '|') {
                    b = q;
                    break;
                }
        }
        for (const char *q = e; q > b; q--)
            if (q[-1] == ':') {
                b = q;
                break;
            }
        const std::string_view ko(b, (size_t)(e - b));
        const int32_t at = index->find(ko.data(), ko.size());
        if (at < 0) {
            res.warnings.append(ko.data(), ko.size());
            res.warnings.push_back('\n');
        } else {
            P[at] = value;
        }
    }
    const double total = np_pairwise_sum(P, n);
    if (total == 0) return set_msg(res, 2, "");
    if (normalize)
        for (int32_t k = 0; k < n; k++) P[k] = P[k] / total;
    if (unweighted)
        for (int32_t k = 0; k < n; k++)
            if (P[k] > 0) P[k] = 1;
}

}  // namespace ingest
}  // namespace fuf

using namespace fuf;

extern "C" int fuf_name_index_create(fuf_name_index **out, const char *const *names, int32_t n)
{
    FUF_REQUIRE(out && names && n >= 0, "fuf_name_index_create: bad argument");
    fuf_name_index *ix = new (std::nothrow) fuf_name_index();
    if (!ix) {
        set_error("fuf_name_index_create: out of host memory");
        return FUF_ENOMEM;
    }
    ix->n = n;
    ix->map.reserve((size_t)n * 2);
    for (int32_t k = 0; k < n; k++) {   // a repeated name keeps its last index (dict semantics)
        ix->names.emplace_back(names[k]);
        ix->map[std::string_view(ix->names.back())] = k;
    }
    // flat table over the final (name -> last index) map: names of 1..7 bytes without NUL bytes
    size_t cap = 64;
    while (cap < (size_t)n * 4) cap <<= 1;
    ix->flat_key.assign(cap, 0);
    ix->flat_val.assign(cap, -1);
    ix->flat_mask = cap - 1;
    for (const auto &kv : ix->map) {
        const std::string_view nm = kv.first;
        if (nm.empty() || nm.size() > 7 || memchr(nm.data(), 0, nm.size())) continue;
        const uint64_t key = fuf_name_index::pack(nm.data(), nm.size());
        uint64_t h = fuf_name_index::mix(key) & ix->flat_mask;
        while (ix->flat_key[h] != 0) h = (h + 1) & ix->flat_mask;
        ix->flat_key[h] = key;
        ix->flat_val[h] = kv.second;
    }
    *out = ix;
    return FUF_OK;
}

extern "C" int fuf_name_index_destroy(fuf_name_index *ix)
{
    delete ix;
    return FUF_OK;
}

extern "C" int fuf_ingest_gather_csv(const fuf_name_index *index, const char *const *paths, int32_t n_files,
                                     const char *abundance_key, int normalize, int unweighted, int n_threads, double *X,
                                     int64_t ldx, int32_t *status, char *messages, int64_t message_stride, char *warnings,
                                     int64_t warnings_cap)
{
    FUF_REQUIRE(index && paths && abundance_key && X && status && n_files >= 0 && ldx >= index->n,
                "fuf_ingest_gather_csv: bad argument");
    FUF_REQUIRE(!messages || message_stride > 0, "fuf_ingest_gather_csv: message_stride must be positive");
    ingest::pow10_table();  // initialise before the threads start
    std::vector<ingest::FileResult> results((size_t)n_files);
    const std::string key(abundance_key);
    std::atomic<int32_t> next(0);
    auto worker = [&]() {
        for (;;) {
            const int32_t f = next.fetch_add(1);
            if (f >= n_files) break;
            ingest::parse_file(index, paths[f], key, normalize != 0, unweighted != 0, X + (size_t)f * ldx, results[f]);
        }
    };
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, std::max(1, (int)n_files)));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(worker);
    worker();
    for (auto &th : pool) th.join();
    int64_t wpos = 0;
    int n_failed = 0;
    for (int32_t f = 0; f < n_files; f++) {
        status[f] = results[f].rc;
        if (results[f].rc) n_failed++;
        if (messages) {
            snprintf(messages + (size_t)f * message_stride, (size_t)message_stride, "%s", results[f].message.c_str());
        }
        if (warnings && !results[f].warnings.empty()) {  // "<file index>\t<ko>\n" per unknown id, in file and row order
            size_t start = 0;
            const std::string &w = results[f].warnings;
            while (start < w.size()) {
                size_t nl = w.find('\n', start);
                std::string line = std::to_string(f) + "\t" + w.substr(start, nl - start) + "\n";
                if (wpos + (int64_t)line.size() < warnings_cap) {
                    memcpy(warnings + wpos, line.data(), line.size());
                    wpos += (int64_t)line.size();
                }
                start = nl + 1;
            }
        }
    }
    if (warnings && warnings_cap > 0) warnings[std::min(wpos, warnings_cap - 1)] = '\0';
    return n_failed ? 1 : FUF_OK;
}
