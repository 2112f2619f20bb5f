// Marker dictionaries of the C ABI: the code tables are the blob tools/extract_dictionaries.py wrote
// (data extracted from /root/reference/src/dictionaries.rs:5-19 and the name map :30-113), embedded at build time.
// Functions restate ARDictionary's methods (/root/reference/src/dictionaries.rs:115-232).
#include <math.h>
#include <string.h>

#include <mutex>

#include "a3_internal.h"

#ifndef A3_DICT_BIN_PATH
#error "define A3_DICT_BIN_PATH (absolute path of aruco3_b200/data/dictionaries.bin)"
#endif
__asm__(".section .rodata\n"
        ".balign 16\n"
        "a3_dict_blob:\n"
        ".incbin \"" A3_DICT_BIN_PATH "\"\n"
        ".byte 0\n"
        ".previous\n");
extern "C" const uint8_t a3_dict_blob[];

namespace {
struct BlobEntry {
    char name[24];
    uint8_t num_bits, tau_table;
    uint16_t reserved;
    uint32_t n_codes, first_code, reserved2;
};
uint32_t n_entries() { uint32_t v; memcpy(&v, a3_dict_blob + 8, 4); return v; }
const BlobEntry *entries() { return reinterpret_cast<const BlobEntry *>(a3_dict_blob + 16); }
const uint64_t *all_codes() { return reinterpret_cast<const uint64_t *>(a3_dict_blob + 16 + sizeof(BlobEntry) * n_entries()); }

// effective tau per entry, computed once (min pairwise distance when the table says 0, dictionaries.rs:124,129-138)
uint8_t effective_tau(uint32_t e) {
    static std::mutex mu;
    static int cache[64];
    static bool init = false;
    std::lock_guard<std::mutex> lock(mu);
    if (!init) { for (int &c : cache) c = -1; init = true; }
    if (e < 64 && cache[e] >= 0) return (uint8_t)cache[e];
    const BlobEntry &be = entries()[e];
    int tau = be.tau_table;
    if (tau == 0) {
        const uint64_t *c = all_codes() + be.first_code;
        tau = 255;
        for (uint32_t i = 0; i < be.n_codes; i++)
            for (uint32_t j = i + 1; j < be.n_codes; j++) {
                const int d = __builtin_popcountll(c[i] ^ c[j]);
                if (d < tau) tau = d;
            }
    }
    if (e < 64) cache[e] = tau;
    return (uint8_t)tau;
}
}  // namespace

namespace a3 {
uint8_t mark_size_of(uint8_t num_bits) { return (uint8_t)((uint8_t)ceilf(sqrtf((float)num_bits)) + 2); }
}  // namespace a3

extern "C" {

int32_t a3_dictionary_count(void) { return (int32_t)n_entries(); }

const char *a3_dictionary_name(int32_t i) {
    if (i < 0 || (uint32_t)i >= n_entries()) return nullptr;
    return entries()[i].name;
}

a3_status a3_dictionary_by_name(const char *name, a3_dictionary *out) {
    if (!name || !out) return a3::fail(A3_ERR_INVALID_ARGUMENT, "a3_dictionary_by_name: null argument");
    char up[24];
    const size_t n = strlen(name);
    if (n >= sizeof(up)) return a3::fail(A3_ERR_UNKNOWN_DICTIONARY, std::string("unknown dictionary: ") + name);
    for (size_t i = 0; i <= n; i++) up[i] = (name[i] >= 'a' && name[i] <= 'z') ? (char)(name[i] - 32) : name[i];
    for (uint32_t e = 0; e < n_entries(); e++) {
        if (strncmp(entries()[e].name, up, sizeof(up)) == 0) {
            out->num_bits = entries()[e].num_bits;
            out->n_codes = entries()[e].n_codes;
            out->codes = all_codes() + entries()[e].first_code;
            out->tau = effective_tau(e);
            return A3_OK;
        }
    }
    return a3::fail(A3_ERR_UNKNOWN_DICTIONARY, std::string("unknown dictionary: ") + name);
}

uint8_t a3_dictionary_mark_size(const a3_dictionary *d) { return d ? a3::mark_size_of(d->num_bits) : 0; }

uint8_t a3_hamming_distance(uint64_t a, uint64_t b) { return (uint8_t)__builtin_popcountll(a ^ b); }

void a3_find_nearest(const a3_dictionary *d, uint64_t bits, uint64_t *index, uint8_t *dist) {
    uint64_t best_i = 0;
    uint8_t best = 0xFF;
    for (uint32_t i = 0; d && i < d->n_codes; i++) {
        const uint8_t dd = (uint8_t)__builtin_popcountll(d->codes[i] ^ bits);
        if (dd < best) { best = dd; best_i = i; }
    }
    if (index) *index = best_i;
    if (dist) *dist = best;
}

int32_t a3_try_find_nearest(const a3_dictionary *d, uint64_t bits, uint64_t *index, uint8_t *dist) {
    uint64_t i; uint8_t dd;
    a3_find_nearest(d, bits, &i, &dd);
    if (index) *index = i;
    if (dist) *dist = dd;
    return d && dd < d->tau;
}

uint8_t a3_make_binary_image(const a3_dictionary *d, uint64_t marker_id, uint8_t *bits, uint32_t capacity, uint32_t *n_bits) {
    if (!d || marker_id >= d->n_codes) { if (n_bits) *n_bits = 0; return 0; }
    const uint64_t code = d->codes[marker_id];
    const uint8_t width = a3::mark_size_of(d->num_bits);
    uint32_t len = 0;
    auto push = [&](uint8_t v) { if (bits && len < capacity) bits[len] = v; len++; };
    for (uint8_t i = 0; i < width; i++) push(0);
    for (uint8_t i = 0; i < d->num_bits; i++) {
        if ((uint8_t)len % width == 0) push(0);
        push((code >> i) & 1);  // least significant bit first: the quirk the reference's own TODO suspects
        if ((uint8_t)len % width == width - 1) push(0);
    }
    for (uint8_t i = 0; i < width; i++) push(0);
    if (n_bits) *n_bits = len;
    return width;
}

}  // extern "C"
