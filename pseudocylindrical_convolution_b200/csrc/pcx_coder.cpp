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

// MSB-first bit output: a 64-bit accumulator, whole bytes moved out when it fills (BitIoStream.cpp:52-72 emits the same bytes)
struct BitSink {
    std::vector<unsigned char> bytes;
    uint64_t acc = 0;
    int nbits = 0;                 // valid low bits of acc, < 8 between calls
    void put_bits(uint32_t v, int n)            // n <= 32, the n low bits of v, most significant first
    {
        if (n <= 0) return;
        acc = (acc << n) | (uint64_t)(n == 32 ? v : (v & ((1u << n) - 1u)));
        nbits += n;
        while (nbits >= 8) {
            nbits -= 8;
            bytes.push_back((unsigned char)(acc >> nbits));
        }
    }
    void put(unsigned bit) { put_bits(bit, 1); }
    void put_run(unsigned bit, uint64_t count)   // `count` copies of one bit
    {
        const uint32_t pat = bit ? 0xffffffffu : 0u;
        while (count > 0) {
            int n = count > 32 ? 32 : (int)count;
            put_bits(pat, n);
            count -= (uint64_t)n;
        }
    }
    void flush() { if (nbits != 0) put_bits(0, 8 - nbits); }
};

// MSB-first bit input; past the end the stream reads as zeros (ArithmeticCoder.cpp:129-134)
struct BitSource {
    std::vector<unsigned char> bytes;
    size_t pos = 0;
    uint64_t acc = 0;
    int left = 0;                  // valid low bits of acc
    uint32_t get_bits(int n)       // n <= 32
    {
        if (n <= 0) return 0;
        while (left < n) {
            acc = (acc << 8) | (pos < bytes.size() ? bytes[pos] : 0u);
            pos++;
            left += 8;
        }
        left -= n;
        return (uint32_t)((acc >> left) & (n == 32 ? 0xffffffffull : ((1ull << n) - 1ull)));
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
                sink.put(bit);
                if (pending > 0) { sink.put_run(bit ^ 1u, pending); pending = 0; }
                if (n > 1) sink.put_bits((uint32_t)(low >> (kStateBits - n)), n - 1);
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

int pcx_coder_encodes(pcx_coder *c, const int32_t *table, int ncode, const int32_t *symbols, int n)
{
    if (!c || !table || !symbols || ncode < 1 || n < 0) { pcx_set_error("coder: bad encodes arguments"); return PCX_EINVAL; }
    if (!c->encoding) { pcx_set_error("coder: encodes before start_encoder"); return PCX_ECODER; }
    const int stride = ncode + 1;
    for (int i = 0; i < n; i++) {
        const uint32_t *cum = reinterpret_cast<const uint32_t *>(table + (size_t)i * stride);
        uint32_t sym = (uint32_t)symbols[i];
        if (sym >= (uint32_t)ncode) { pcx_set_error("coder: symbol %u outside the %d-entry table", sym, ncode); return PCX_ECODER; }
        int rc = c->narrow<true>(cum, cum[ncode], sym);
        if (rc) return rc;
    }
    return PCX_OK;
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
    c->source.pos = 0;
    c->source.left = 0;
    c->source.acc = 0;
    c->code = c->source.get_bits(kStateBits);
    return PCX_OK;
}

int pcx_coder_start_decoder_mem(pcx_coder *c, const unsigned char *src, long long n)
{
    if (!c || (!src && n > 0) || n < 0) return PCX_EINVAL;
    c->source = BitSource();
    c->source.bytes.assign(src, src + n);
    return begin_decode(c);
}

int pcx_coder_start_decoder(pcx_coder *c)
{
    if (!c) return PCX_EINVAL;
    FILE *f = fopen(c->path.c_str(), "rb");
    if (!f) { pcx_set_error("coder: cannot open %s for reading", c->path.c_str()); return PCX_EIO; }
    c->source = BitSource();
    unsigned char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) c->source.bytes.insert(c->source.bytes.end(), buf, buf + got);
    fclose(f);
    return begin_decode(c);
}

int pcx_coder_decodes(pcx_coder *c, const int32_t *table, int ncode, int n, float *out_symbols)
{
    if (!c || !table || !out_symbols || ncode < 1 || n < 0) { pcx_set_error("coder: bad decodes arguments"); return PCX_EINVAL; }
    const int stride = ncode + 1;
    for (int i = 0; i < n; i++) {
        const uint32_t *cum = reinterpret_cast<const uint32_t *>(table + (size_t)i * stride);
        uint32_t sym = 0;
        int rc = c->decode_one(cum, (uint32_t)ncode, cum[ncode], &sym);
        if (rc) return rc;
        out_symbols[i] = (float)sym;
    }
    return PCX_OK;
}

}  // extern "C"
