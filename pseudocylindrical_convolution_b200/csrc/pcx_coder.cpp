// pcx_coder.cpp - host range coder behind the `coder.coder` interface (coder/python.cpp:63-72).
//
// Wire format: 32-bit-state arithmetic coder, cumulative-frequency tables of ncode+1
// uint32 entries per symbol with total = table[ncode], bits emitted MSB first, no header, one terminating
// `1` bit, zero padding to the byte (coder/ArithmeticCoder.cpp:34-69, :82-116, :152-154;
// coder/BitIoStream.cpp:52-72).  The serial coder stays on host threads by design (north_star); this
// implementation keeps the whole bitstream in memory so it can run next to the GPU pipeline without
// per-symbol stream I/O, and touches the file only in start_decoder / end_encoder.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <emmintrin.h>
#include "../../include/pcx.h"

void pcx_set_error(const char *fmt, ...);

namespace {

constexpr int kStateBits = 32;
constexpr uint64_t kFull = 1ull << kStateBits;       // 2^32
constexpr uint64_t kMask = kFull - 1;
constexpr uint64_t kHalf = kFull >> 1;               // top bit
constexpr uint64_t kQuarter = kHalf >> 1;            // second bit
constexpr uint64_t kMinRange = (kFull >> 2) + 2;
constexpr uint64_t kMaxTotal = kMinRange;            // min(2^64/2^32, MIN_RANGE)

// The coder's per-symbol work is ~40 arithmetic instructions; what made it 25-45 ns per symbol were its DATA-DEPENDENT
// branches (how many bits leave, whether underflow bits are pending, the decoder's search): one misprediction costs more
// than the arithmetic.  Bit I/O below is therefore branch-free: every call moves a 64-bit window to / from memory.
//
// MSB-first bit output (BitIoStream.cpp:52-72 emits the same bytes one bit at a time).  `bytes` is kept over-allocated while
// encoding; `len` counts the bytes that are final and flush() trims the vector to it.
struct BitSink {
    std::vector<unsigned char> bytes;
    size_t len = 0;
    uint64_t acc = 0;
    int nbits = 0;                 // bits of acc not yet final in `bytes`, < 8 between calls
    void reserve(size_t more)      // room for `more` further bytes (+ the 8-byte store window)
    {
        if (len + more + 16 > bytes.size()) bytes.resize((len + more + 16) * 2 + 4096);
    }
    inline void put_bits(uint32_t v, int n)     // n in [0, 32]: the n low bits of v, most significant first; caller reserved the room
    {
        acc = (acc << n) | ((uint64_t)v & ((1ull << n) - 1ull));
        nbits += n;                             // <= 39
        const uint64_t w = __builtin_bswap64((acc << 1) << (63 - nbits));    // pending bits left-aligned (nbits = 0: one junk byte, rewritten later)
        memcpy(bytes.data() + len, &w, 8);
        len += (size_t)(nbits >> 3);
        nbits &= 7;
    }
    void put(unsigned bit) { reserve(8); put_bits(bit, 1); }
    void put_run(unsigned bit, uint64_t count)   // `count` copies of one bit
    {
        reserve((size_t)(count >> 3) + 8);
        const uint32_t pat = bit ? 0xffffffffu : 0u;
        while (count > 0) {
            int n = count > 32 ? 32 : (int)count;
            put_bits(pat, n);
            count -= (uint64_t)n;
        }
    }
    void flush()                   // zero padding to the byte, then the exact byte count
    {
        reserve(8);
        if (nbits > 0) bytes[len++] = (unsigned char)((acc << (8 - nbits)) & 0xffu);
        nbits = 0;
        bytes.resize(len);
    }
};

// MSB-first bit input; past the end the stream reads as zeros (ArithmeticCoder.cpp:129-134): `bytes` carries 16 zero bytes of
// padding behind the `nbytes` of the stream and the read window is clamped into it.
struct BitSource {
    std::vector<unsigned char> bytes;
    size_t nbytes = 0;
    uint64_t bitpos = 0;
    void assign(const unsigned char *src, size_t n)
    {
        bytes.assign(n + 16, 0);
        if (n) memcpy(bytes.data(), src, n);
        nbytes = n;
        bitpos = 0;
    }
    inline uint32_t get_bits(int n)       // n in [0, 32]
    {
        size_t byte = (size_t)(bitpos >> 3);
        byte = byte < nbytes + 8 ? byte : nbytes + 8;
        const int s = (int)(bitpos & 7);
        uint64_t w;
        memcpy(&w, bytes.data() + byte, 8);
        w = __builtin_bswap64(w);
        bitpos += (uint64_t)n;
        return (uint32_t)((((w << s) >> 1)) >> (63 - n));
    }
    unsigned get() { return get_bits(1); }
};

}  // namespace

struct pcx_coder {
    std::string path;
    uint64_t low = 0, high = kMask, code = 0;
    uint64_t pending = 0;      // underflow bits waiting for the next shifted bit
    bool encoding = false, to_file = true;
    BitSink sink;
    BitSource source;

    void reset() { low = 0; high = kMask; code = 0; pending = 0; }

    // narrows [low, high] to the symbol's sub-range and renormalises; emit/consume bits through F
    template <bool ENC>
    int narrow(const uint32_t *cum, uint32_t total, uint32_t sym)
    {
        if (low >= high || (low & kMask) != low || (high & kMask) != high) { pcx_set_error("coder: low/high out of range"); return PCX_ECODER; }
        const uint64_t range = high - low + 1;
        if (range < kMinRange || range > kFull) { pcx_set_error("coder: range out of range"); return PCX_ECODER; }
        const uint32_t lo = cum[sym], hi = cum[sym + 1];
        if (lo == hi) { pcx_set_error("coder: symbol %u has zero frequency", sym); return PCX_ECODER; }
        if (total > kMaxTotal) { pcx_set_error("coder: total %u too large", total); return PCX_ECODER; }
        // total is a power of two for every table the GMM stage emits (65536): the exact division becomes a shift
        const bool p2 = (total & (total - 1)) == 0;
        const int sh = __builtin_ctz(total);
        const uint64_t nl = low + (p2 ? ((uint64_t)lo * range) >> sh : (uint64_t)lo * range / total);
        const uint64_t nh = low + (p2 ? ((uint64_t)hi * range) >> sh : (uint64_t)hi * range / total) - 1;
        low = nl;
        high = nh;
        // The reference shifts one bit per iteration (ArithmeticCoder.cpp:51-68); here all the leading bits on which low and
        // high agree leave in one go, then all the underflow bits (low = 01.., high = 10..) - the same bits in the same order.
        const uint32_t diff = (uint32_t)(low ^ high);
        const int n = diff == 0 ? 32 : __builtin_clz(diff);
        if (n > 0) {
            if (ENC) {
                const unsigned bit = (unsigned)(low >> (kStateBits - 1));
                if (pending > 0) {
                    sink.put(bit);
                    sink.put_run(bit ^ 1u, pending);
                    pending = 0;
                    sink.reserve(8);
                    sink.put_bits((uint32_t)(low >> (kStateBits - n)), n - 1);
                } else {
                    sink.reserve(8);
                    sink.put_bits((uint32_t)(low >> (kStateBits - n)), n);      // the leading bit and the n - 1 that follow it
                }
            } else {
                code = n == 32 ? source.get_bits(32) : (((code << n) & kMask) | source.get_bits(n));
            }
            low = n == 32 ? 0 : (low << n) & kMask;
            high = n == 32 ? kMask : (((high << n) & kMask) | ((1ull << n) - 1ull));
        }
        const uint32_t under = (uint32_t)(low & ~high) << 1;   // bit 31 <- bit 30: run of positions with low = 1, high = 0
        const int m = under == 0xffffffffu ? 31 : __builtin_clz(~under);
        if (m > 0) {
            if (ENC) pending += (uint64_t)m;
            else code = (code & kHalf) | ((code << m) & (kMask >> 1)) | source.get_bits(m);
            low = (low << m) & (kMask >> 1);
            high = ((high << m) & (kMask >> 1)) | kHalf | ((1ull << m) - 1ull);
        }
        return PCX_OK;
    }

    int decode_one(const uint32_t *cum, uint32_t ncode, uint32_t total, uint32_t *out)
    {
        if (total > kMaxTotal) { pcx_set_error("coder: total %u too large", total); return PCX_ECODER; }
        const uint64_t range = high - low + 1;
        const uint64_t offset = code - low;
        const uint64_t value = ((offset + 1) * total - 1) / range;
        const bool p2 = (total & (total - 1)) == 0;
        const int sh = __builtin_ctz(total);
        if ((p2 ? (value * range) >> sh : value * range / total) > offset || value >= total) { pcx_set_error("coder: decoder state inconsistent"); return PCX_ECODER; }
        uint32_t a = 0, b = ncode;                        // highest symbol with cum[symbol] <= value
        while (b - a > 1) {
            uint32_t mid = (a + b) >> 1;
            if (cum[mid] > value) b = mid; else a = mid;
        }
        if (offset < (p2 ? ((uint64_t)cum[a] * range) >> sh : (uint64_t)cum[a] * range / total) ||
            (p2 ? ((uint64_t)cum[a + 1] * range) >> sh : (uint64_t)cum[a + 1] * range / total) <= offset) {
            pcx_set_error("coder: table does not bracket the code value (encoder/decoder CDF mismatch?)");
            return PCX_ECODER;
        }
        int rc = narrow<false>(cum, total, a);
        if (rc) return rc;
        if (code < low || code > high) { pcx_set_error("coder: code out of range"); return PCX_ECODER; }
        *out = a;
        return PCX_OK;
    }
};

extern "C" {

pcx_coder *pcx_coder_open(const char *path)
{
    pcx_coder *c = new pcx_coder();
    c->path = path ? path : "";
    return c;
}

void pcx_coder_close(pcx_coder *c) { delete c; }

int pcx_coder_start_encoder_mem(pcx_coder *c)
{
    if (!c) return PCX_EINVAL;
    c->reset();
    c->sink = BitSink();
    c->encoding = true;
    c->to_file = false;
    return PCX_OK;
}

int pcx_coder_start_encoder(pcx_coder *c)
{
    int rc = pcx_coder_start_encoder_mem(c);
    if (rc) return rc;
    c->to_file = true;
    FILE *f = fopen(c->path.c_str(), "wb");          // the reference opens (truncates) the file here (coder.h:14-20)
    if (!f) { pcx_set_error("coder: cannot open %s for writing", c->path.c_str()); return PCX_EIO; }
    fclose(f);
    return PCX_OK;
}

// Batched forms of narrow<>: the coder state lives in locals for the whole span (the member form reloads it around every byte
// store), the invariants that cannot fail by construction are checked once on entry, and the decoder finds its symbol by a
// binary search over the scaled boundaries (cum[j] * range) >> log2(total) instead of the reference's 64-bit division - the
// symbol is the unique j with boundary(j) <= offset < boundary(j + 1), which is exactly the bracket the division-based form
// verifies after the fact (ArithmeticCoder.cpp:93-107), so streams, symbols and error cases are unchanged.
int pcx_coder_encodes(pcx_coder *c, const int32_t *table, int ncode, const int32_t *symbols, int n)
{
    if (!c || !table || !symbols || ncode < 1 || n < 0) { pcx_set_error("coder: bad encodes arguments"); return PCX_EINVAL; }
    if (!c->encoding) { pcx_set_error("coder: encodes before start_encoder"); return PCX_ECODER; }
    uint64_t low = c->low, high = c->high, pending = c->pending;
    if (low >= high || (low & kMask) != low || (high & kMask) != high || high - low + 1 < kMinRange) { pcx_set_error("coder: low/high out of range"); return PCX_ECODER; }
    BitSink &sink = c->sink;
    const int stride = ncode + 1;
    int rc = PCX_OK;
    // room for the whole span up front: a symbol emits at most 31 fresh bits (k <= 31) and the pending run is reserved where it
    // is flushed, so the common path below never has to check the buffer
    sink.reserve((size_t)n * 4 + 16);
    // the sink's cursor lives in locals too (a byte store through the vector's pointer would otherwise force reloads of its
    // members); the rare long-run path below goes through the member functions and re-reads them afterwards
    unsigned char *base = sink.bytes.data();
    size_t len = sink.len;
    uint64_t acc = sink.acc;
    int nbits = sink.nbits;
    for (int i = 0; i < n; i++) {
        const uint32_t *cum = reinterpret_cast<const uint32_t *>(table + (size_t)i * stride);
        const uint32_t sym = (uint32_t)symbols[i];
        if (sym >= (uint32_t)ncode) { pcx_set_error("coder: symbol %u outside the %d-entry table", sym, ncode); rc = PCX_ECODER; break; }
        const uint32_t total = cum[ncode], lo = cum[sym], hi = cum[sym + 1];
        if (lo == hi) { pcx_set_error("coder: symbol %u has zero frequency", sym); rc = PCX_ECODER; break; }
        if (total > kMaxTotal) { pcx_set_error("coder: total %u too large", total); rc = PCX_ECODER; break; }
        const uint64_t range = high - low + 1;
        uint64_t nl, nh;
        if ((total & (total - 1)) == 0) {                  // every table the GMM stage emits: total = 65536
            const int sh = __builtin_ctz(total);
            nl = low + (((uint64_t)lo * range) >> sh);
            nh = low + (((uint64_t)hi * range) >> sh) - 1;
        } else {
            nl = low + (uint64_t)lo * range / total;
            nh = low + (uint64_t)hi * range / total - 1;
        }
        low = nl;
        high = nh;
        // Leading bits on which low and high agree leave at once (the reference shifts one per iteration,
        // ArithmeticCoder.cpp:51-68), then the underflow bits are counted.  range >= 2^30 and total <= 2^30 + 2 keep low < high,
        // so k <= 31; k = 0 and m = 0 make every expression below an identity - no data-dependent branch except the rare
        // "pending bits and a long run" case.
        const int k = __builtin_clz((uint32_t)(low ^ high) | 1u);      // low != high: the | 1 only guards clz(0)
        const uint32_t head = (uint32_t)((low >> 1) >> (kStateBits - 1 - k));     // the k leading bits of low (k = 0: none)
        if (__builtin_expect(pending + (uint64_t)k <= 32, 1)) {
            // first bit b, then `pending` copies of !b, then the other k - 1 bits; nothing at all when k = 0
            const uint32_t b = (uint32_t)(low >> (kStateBits - 1)) & 1u;
            const int p = k > 0 ? (int)pending : 0;
            const uint32_t rest = k > 0 ? head & ((1u << (k - 1)) - 1u) : 0u;
            const uint64_t fill = b ? 0ull : ((1ull << p) - 1ull);
            const uint64_t word = k > 0 ? ((((uint64_t)b << p) | fill) << (k - 1)) | rest : 0ull;
            const int nb = k + p;                                      // BitSink::put_bits on the local cursor
            acc = (acc << nb) | (word & ((1ull << nb) - 1ull));
            nbits += nb;
            const uint64_t w = __builtin_bswap64((acc << 1) << (63 - nbits));
            memcpy(base + len, &w, 8);
            len += (size_t)(nbits >> 3);
            nbits &= 7;
            pending -= (uint64_t)p;
        } else if (k > 0) {
            sink.len = len; sink.acc = acc; sink.nbits = nbits;
            const unsigned bit = (unsigned)(low >> (kStateBits - 1));
            sink.put(bit);
            sink.put_run(bit ^ 1u, pending);
            pending = 0;
            sink.reserve((size_t)(n - i) * 4 + 16);
            sink.put_bits(head, k - 1);
            base = sink.bytes.data(); len = sink.len; acc = sink.acc; nbits = sink.nbits;
        }
        low = (low << k) & kMask;
        high = ((high << k) & kMask) | ((1ull << k) - 1ull);
        const uint32_t under = (uint32_t)(low & ~high) << 1;           // bit 31 <- bit 30: run of positions with low = 1, high = 0
        const int m = under == 0xffffffffu ? 31 : __builtin_clz(~under);
        pending += (uint64_t)m;
        low = (low << m) & (kMask >> 1);                               // bit 31 of low is 0 and of high is 1 here: identities for m = 0
        high = ((high << m) & (kMask >> 1)) | kHalf | ((1ull << m) - 1ull);
    }
    sink.len = len; sink.acc = acc; sink.nbits = nbits;
    c->low = low; c->high = high; c->pending = pending;
    return rc;
}

int pcx_coder_end_encoder(pcx_coder *c)
{
    if (!c || !c->encoding) { pcx_set_error("coder: end_encoder without start_encoder"); return PCX_ECODER; }
    c->sink.put(1);                                   // ArithmeticEncoder::finish
    c->sink.flush();                                  // BitOutputStream::finish
    c->encoding = false;
    if (c->to_file) {
        FILE *f = fopen(c->path.c_str(), "wb");
        if (!f) { pcx_set_error("coder: cannot open %s for writing", c->path.c_str()); return PCX_EIO; }
        size_t n = c->sink.bytes.size();
        size_t w = n ? fwrite(c->sink.bytes.data(), 1, n, f) : 0;
        fclose(f);
        if (w != n) { pcx_set_error("coder: short write to %s", c->path.c_str()); return PCX_EIO; }
    }
    return PCX_OK;
}

long long pcx_coder_take_bytes(pcx_coder *c, unsigned char *dst, long long cap)
{
    if (!c) return PCX_EINVAL;
    long long n = (long long)c->sink.bytes.size();
    if (dst) {
        if (cap < n) { pcx_set_error("coder: buffer of %lld bytes too small for %lld", cap, n); return PCX_EINVAL; }
        if (n) memcpy(dst, c->sink.bytes.data(), (size_t)n);
    }
    return n;
}

static int begin_decode(pcx_coder *c)
{
    c->reset();
    c->encoding = false;
    c->source.bitpos = 0;
    c->code = c->source.get_bits(kStateBits);
    return PCX_OK;
}

int pcx_coder_start_decoder_mem(pcx_coder *c, const unsigned char *src, long long n)
{
    if (!c || (!src && n > 0) || n < 0) return PCX_EINVAL;
    c->source = BitSource();
    c->source.assign(src, (size_t)n);
    return begin_decode(c);
}

int pcx_coder_start_decoder(pcx_coder *c)
{
    if (!c) return PCX_EINVAL;
    FILE *f = fopen(c->path.c_str(), "rb");
    if (!f) { pcx_set_error("coder: cannot open %s for reading", c->path.c_str()); return PCX_EIO; }
    c->source = BitSource();
    std::vector<unsigned char> all;
    unsigned char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) all.insert(all.end(), buf, buf + got);
    fclose(f);
    c->source.assign(all.data(), all.size());
    return begin_decode(c);
}

int pcx_coder_decodes(pcx_coder *c, const int32_t *table, int ncode, int n, float *out_symbols)
{
    if (!c || !table || !out_symbols || ncode < 1 || n < 0) { pcx_set_error("coder: bad decodes arguments"); return PCX_EINVAL; }
    uint64_t low = c->low, high = c->high, code = c->code;
    if (low >= high || (low & kMask) != low || (high & kMask) != high || high - low + 1 < kMinRange) { pcx_set_error("coder: low/high out of range"); return PCX_ECODER; }
    BitSource &src = c->source;
    const int stride = ncode + 1;
    int rc = PCX_OK;
    for (int i = 0; i < n; i++) {
        const uint32_t *cum = reinterpret_cast<const uint32_t *>(table + (size_t)i * stride);
        const uint32_t total = cum[ncode];
        if (total > kMaxTotal || total == 0) { pcx_set_error("coder: total %u too large", total); rc = PCX_ECODER; break; }
        const uint64_t range = high - low + 1, offset = code - low;
        const bool p2 = (total & (total - 1)) == 0;
        const int sh = p2 ? __builtin_ctz(total) : 0;
        // boundaries of all symbols at once and a branch-free count of those at or below the offset (ncode = 8 in this codec)
        uint64_t bnd[40];
        uint32_t a;
        if (__builtin_expect(ncode <= 32, 1)) {
            bnd[0] = p2 ? ((uint64_t)cum[0] * range) >> sh : (uint64_t)cum[0] * range / total;
            uint32_t cnt = 0;
            for (int j = 1; j <= ncode; j++) {
                bnd[j] = p2 ? ((uint64_t)cum[j] * range) >> sh : (uint64_t)cum[j] * range / total;
                cnt += (j < ncode && bnd[j] <= offset) ? 1u : 0u;
            }
            a = cnt;                                            // boundaries are non-decreasing: the count is the highest such index
        } else {
            pcx_set_error("coder: batched decoding supports at most 32 symbols per table (got %d)", ncode);
            rc = PCX_EINVAL;
            break;
        }
        const uint64_t ba = bnd[a], bb = bnd[a + 1];
        if (code < low || offset < ba || bb <= offset) {
            pcx_set_error("coder: table does not bracket the code value (encoder/decoder CDF mismatch?)");
            rc = PCX_ECODER;
            break;
        }
        high = low + bb - 1;
        low = low + ba;
        const int k = __builtin_clz((uint32_t)(low ^ high) | 1u);      // see pcx_coder_encodes: k <= 31, k = 0 / m = 0 are identities
        code = ((code << k) & kMask) | src.get_bits(k);
        low = (low << k) & kMask;
        high = ((high << k) & kMask) | ((1ull << k) - 1ull);
        const uint32_t under = (uint32_t)(low & ~high) << 1;
        const int m = under == 0xffffffffu ? 31 : __builtin_clz(~under);
        code = (code & kHalf) | ((code << m) & (kMask >> 1)) | src.get_bits(m);
        low = (low << m) & (kMask >> 1);
        high = ((high << m) & (kMask >> 1)) | kHalf | ((1ull << m) - 1ull);
        out_symbols[i] = (float)a;
    }
    c->low = low; c->high = high; c->code = code;
    return rc;
}

// The loop of pcx_coder_decodes_rows16, compiled twice: for the baseline x86-64 and for BMI2 / LZCNT (three-operand shifts by a
// register, lzcnt instead of bsr + xor, andn) - about 150 vs 115 instructions per symbol, all on the serial range -> symbol ->
// renormalisation chain that the device waits for.  Picked once per process by cpuid.
static inline __attribute__((always_inline)) int rows16_loop(pcx_coder *c, const uint16_t *rows, int n, unsigned tag16, unsigned word_tag,
                                                             uint32_t *out_words, int *done)
{
    uint64_t low = c->low, high = c->high, code = c->code;
    if (low >= high || (low & kMask) != low || (high & kMask) != high || high - low + 1 < kMinRange) { pcx_set_error("coder: low/high out of range"); return PCX_ECODER; }
    BitSource &src = c->source;
    // the stream cursor lives in locals; every symbol consumes k + m bits (k leading agreeing bits, m underflow bits), both taken
    // from ONE 64-bit window fetched at the top of the iteration - its address depends only on the cursor, so the load is off
    // the serial chain range -> boundaries -> symbol -> low / high -> k -> m
    const unsigned char *bytes = src.bytes.data();
    const size_t last_window = src.nbytes + 8;                     // the vector carries 16 zero bytes behind the stream
    uint64_t bitpos = src.bitpos;
    int rc = PCX_OK, i = 0;
    for (; i < n; i++) {
        // ONE 16-byte load, forced (a plain _mm_load_si128 is an ordinary dereference: the optimiser narrowed it into separate
        // loads and read the boundaries BEFORE the tag - a row landing in between was accepted with the previous step's
        // boundaries).  The device wrote the row with a single 16-byte store, so tag and boundaries are one snapshot.
        __m128i v;
        __asm__ __volatile__("movdqa %1, %0" : "=x"(v) : "m"(*reinterpret_cast<const __m128i *>(rows + (size_t)i * 8)) : "memory");
        uint16_t r[8];
        memcpy(r, &v, 16);
        if (r[7] != (uint16_t)tag16) break;
        size_t byte = (size_t)(bitpos >> 3);
        byte = byte < last_window ? byte : last_window;
        uint64_t win;
        memcpy(&win, bytes + byte, 8);
        win = __builtin_bswap64(win) << (bitpos & 7);              // next 57 .. 64 bits of the stream, left-aligned
        const uint64_t range = high - low + 1, offset = code - low;
        uint64_t bnd[9];
        bnd[0] = 0;
        uint32_t a = 0;
        for (int j = 1; j < 8; j++) {
            bnd[j] = ((uint64_t)r[j - 1] * range) >> 16;
            a += bnd[j] <= offset ? 1u : 0u;               // boundaries are non-decreasing: the count is the symbol
        }
        bnd[8] = range;                                    // (65536 * range) >> 16
        const uint64_t ba = bnd[a], bb = bnd[a + 1];
        if (code < low || offset < ba || bb <= offset) {
            pcx_set_error("coder: table does not bracket the code value (encoder/decoder CDF mismatch?)");
            rc = PCX_ECODER;
            break;
        }
        high = low + bb - 1;
        low = low + ba;
        const int k = __builtin_clz((uint32_t)(low ^ high) | 1u);      // see pcx_coder_encodes: k <= 31, k = 0 / m = 0 are identities
        low = (low << k) & kMask;
        high = ((high << k) & kMask) | ((1ull << k) - 1ull);
        const uint32_t under = (uint32_t)(low & ~high) << 1;
        const int m = under == 0xffffffffu ? 31 : __builtin_clz(~under);
        uint64_t bits_k, bits_m;
        if (__builtin_expect(k + m <= 56, 1)) {
            bits_k = (win >> 1) >> (63 - k);                       // top k bits of the window (k = 0: none)
            bits_m = ((win << k) >> 1) >> (63 - m);                // the next m bits
        } else {                                                   // more than one window's worth: fetch the m bits separately
            bits_k = (win >> 1) >> (63 - k);
            size_t byte2 = (size_t)((bitpos + (uint64_t)k) >> 3);
            byte2 = byte2 < last_window ? byte2 : last_window;
            uint64_t w2;
            memcpy(&w2, bytes + byte2, 8);
            w2 = __builtin_bswap64(w2) << ((bitpos + (uint64_t)k) & 7);
            bits_m = (w2 >> 1) >> (63 - m);
        }
        bitpos += (uint64_t)(k + m);
        code = ((code << k) & kMask) | bits_k;
        code = (code & kHalf) | ((code << m) & (kMask >> 1)) | bits_m;
        low = (low << m) & (kMask >> 1);
        high = ((high << m) & (kMask >> 1)) | kHalf | ((1ull << m) - 1ull);
        __atomic_store_n(out_words + i, (word_tag << 8) | a, __ATOMIC_RELEASE);
    }
    src.bitpos = bitpos;
    c->low = low; c->high = high; c->code = code;
    *done = i;
    return rc;
}
static int rows16_generic(pcx_coder *c, const uint16_t *rows, int n, unsigned tag16, unsigned word_tag, uint32_t *out_words, int *done)
{
    return rows16_loop(c, rows, n, tag16, word_tag, out_words, done);
}
__attribute__((target("bmi,bmi2,lzcnt"))) static int rows16_bmi2(pcx_coder *c, const uint16_t *rows, int n, unsigned tag16, unsigned word_tag,
                                                                 uint32_t *out_words, int *done)
{
    return rows16_loop(c, rows, n, tag16, word_tag, out_words, done);
}

// Decoder side of the persistent wavefront kernel (pcx_flow.cu): the device writes one 16-byte row per symbol into mapped
// pinned memory - cum[1..7] as uint16 and a 16-bit tag in the last half-word (cum[0] = 0 and cum[8] = 65536 are constants of
// the GMM stage) - and this function decodes rows [0, n) for as long as their tags equal `tag16`, i.e. as far as the device
// has got.  Each symbol goes back as one 32-bit word (word_tag << 8 | symbol) that the device polls.  The arithmetic is
// pcx_coder_decodes with total = 2^16; *done = rows consumed (0 when the first row is not there yet).
int pcx_coder_decodes_rows16(pcx_coder *c, const uint16_t *rows, int n, unsigned tag16, unsigned word_tag, uint32_t *out_words, int *done)
{
    if (!c || !rows || !out_words || !done || n < 0) { pcx_set_error("coder: bad decodes_rows16 arguments"); return PCX_EINVAL; }
    *done = 0;
    typedef int (*Fn)(pcx_coder *, const uint16_t *, int, unsigned, unsigned, uint32_t *, int *);
    static const Fn fn = (__builtin_cpu_supports("bmi2") && __builtin_cpu_supports("bmi") && __builtin_cpu_supports("lzcnt") && !getenv("PCX_CODER_GENERIC"))
                             ? rows16_bmi2 : rows16_generic;
    return fn(c, rows, n, tag16, word_tag, out_words, done);
}

}  // extern "C"
