// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library.
//
// Iteration-order emulator for Rust's std::collections::HashMap/HashSet (hashbrown
// SwissTable) with the `fnv` crate's FNV-1a hasher.  The reference's read_id results
// depend on the iteration order of `FnvHashSet<String>` (read_id_mt_pe.rs:115 `for k in map`,
// built by kmer.rs:221-243) and of `FnvHashMap<usize,usize>` (read_id_mt_pe.rs:195
// `report.iter().collect()` before a stable sort).  Neither crate is vendored under
// /root/reference (Cargo.toml:17 `fnv = "1.0.6"`; hashbrown comes with std), so this is a
// restatement of their published behaviour: PARITY UNPINNED.
//
// Modelled literally (control bytes, 16-wide SSE2 groups, trailing mirror bytes) so that it
// is an independent check of the product's bitmap-based device emulation.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <cstring>

namespace orc {

// fnv crate: FnvHasher::default() = 0xcbf29ce484222325; write(): h ^= b; h *= 0x100000001b3
static inline uint64_t fnv1a_bytes(uint64_t h, const uint8_t* p, size_t n) {
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return h;
}
static const uint64_t kFnvBasis = 0xcbf29ce484222325ULL;
// impl Hash for str: write(bytes) then write_u8(0xff)
static inline uint64_t fnv_hash_str(const std::string& s) {
    uint64_t h = fnv1a_bytes(kFnvBasis, (const uint8_t*)s.data(), s.size());
    uint8_t ff = 0xff;
    return fnv1a_bytes(h, &ff, 1);
}
// impl Hash for usize: write(&i.to_ne_bytes()) (little-endian on x86_64)
static inline uint64_t fnv_hash_usize(uint64_t v) {
    uint8_t b[8]; memcpy(b, &v, 8);
    return fnv1a_bytes(kFnvBasis, b, 8);
}

struct HbPolicy {
    int group_width = 16;          // x86_64 SSE2; aarch64 NEON uses 8
    bool reserve_before_find = true;  // hashbrown >= 0.14 HashMap::insert / HashSet::insert
};

// Key-agnostic SwissTable shape emulator. Slots hold an index into the caller's key array.
class HbTable {
public:
    explicit HbTable(HbPolicy pol = HbPolicy()) : pol_(pol) {}

    size_t buckets() const { return buckets_; }
    size_t items() const { return items_; }

    // HashSet::insert / HashMap::insert: (modern) reserve(1) first, then find, then insert.
    // `eq(slot_key_index)` answers key equality.  Returns true if newly inserted.
    template <class Eq>
    bool insert(uint64_t hash, int key_index, Eq eq) {
        if (pol_.reserve_before_find) {
            reserve1();
            if (find(hash, eq) >= 0) return false;
        } else {
            if (find(hash, eq) >= 0) return false;
            reserve1();
        }
        insert_no_grow(hash, key_index);
        return true;
    }
    // HashMap::entry(k).or_insert(v): find first; reserve(1) only when vacant (rustc_entry).
    template <class Eq>
    bool entry_or_insert(uint64_t hash, int key_index, Eq eq) {
        if (find(hash, eq) >= 0) return false;
        reserve1();
        insert_no_grow(hash, key_index);
        return true;
    }
    // Iteration = ascending bucket index over full buckets.
    std::vector<int> iter_order() const {
        std::vector<int> out;
        for (size_t i = 0; i < buckets_; i++) if (!(ctrl_[i] & 0x80)) out.push_back(slot_key_[i]);
        return out;
    }
    template <class Eq>
    int find(uint64_t hash, Eq eq) const {
        if (buckets_ == 0) return -1;
        const size_t mask = buckets_ - 1;
        const uint8_t h2 = (uint8_t)(hash >> 57);
        size_t pos = (size_t)hash & mask, stride = 0;
        for (;;) {
            bool any_empty = false;
            for (int b = 0; b < pol_.group_width; b++) {
                uint8_t c = ctrl_[pos + b];
                if (c == h2) {
                    size_t idx = (pos + b) & mask;
                    if (eq(slot_key_[idx])) return (int)idx;
                }
                if (c == 0xFF) any_empty = true;
            }
            if (any_empty) return -1;
            stride += pol_.group_width;
            pos = (pos + stride) & mask;
        }
    }

private:
    HbPolicy pol_;
    size_t buckets_ = 0;        // 0 = the static empty singleton
    size_t items_ = 0, growth_left_ = 0;
    std::vector<uint8_t> ctrl_;  // buckets_ + group_width bytes; 0xFF = EMPTY
    std::vector<int> slot_key_;
    std::vector<uint64_t> slot_hash_;

    static size_t cap_of(size_t buckets) {  // bucket_mask_to_capacity
        if (buckets == 0) return 0;
        return buckets < 8 ? buckets - 1 : buckets / 8 * 7;
    }
    static size_t capacity_to_buckets(size_t cap) {
        if (cap < 8) return cap < 4 ? 4 : 8;
        size_t adj = cap * 8 / 7, b = 1;
        while (b < adj) b <<= 1;
        return b;
    }
    void reserve1() {
        if (growth_left_ >= 1) return;
        // reserve_rehash: no tombstones ever exist here (no removals), so always resize
        size_t full_cap = cap_of(buckets_);
        size_t want = items_ + 1 > full_cap + 1 ? items_ + 1 : full_cap + 1;
        resize(capacity_to_buckets(want));
    }
    void set_ctrl(size_t i, uint8_t c) {
        const size_t mask = buckets_ - 1, gw = (size_t)pol_.group_width;
        size_t i2 = ((i - gw) & mask) + gw;  // wrapping_sub, as in hashbrown
        ctrl_[i] = c;
        ctrl_[i2] = c;
    }
    size_t find_insert_slot(uint64_t hash) const {
        const size_t mask = buckets_ - 1;
        size_t pos = (size_t)hash & mask, stride = 0;
        for (;;) {
            for (int b = 0; b < pol_.group_width; b++) {
                if (ctrl_[pos + b] & 0x80) {  // EMPTY or DELETED
                    size_t idx = (pos + b) & mask;
                    if (!(ctrl_[idx] & 0x80)) {
                        // fix_insert_slot: small table, hit a trailing fake EMPTY byte
                        for (int g = 0; g < pol_.group_width; g++)
                            if (ctrl_[g] & 0x80) return (size_t)g;
                    }
                    return idx;
                }
            }
            stride += pol_.group_width;
            pos = (pos + stride) & mask;
        }
    }
    void insert_no_grow(uint64_t hash, int key_index) {
        size_t idx = find_insert_slot(hash);
        uint8_t old = ctrl_[idx];
        growth_left_ -= (old & 1);  // EMPTY (0xFF) consumes growth; DELETED (0x80) does not
        set_ctrl(idx, (uint8_t)(hash >> 57));
        slot_key_[idx] = key_index;
        slot_hash_[idx] = hash;
        items_++;
    }
    void resize(size_t new_buckets) {
        std::vector<uint8_t> octrl; octrl.swap(ctrl_);
        std::vector<int> okey; okey.swap(slot_key_);
        std::vector<uint64_t> ohash; ohash.swap(slot_hash_);
        size_t ob = buckets_;
        buckets_ = new_buckets;
        ctrl_.assign(new_buckets + pol_.group_width, 0xFF);
        slot_key_.assign(new_buckets, -1);
        slot_hash_.assign(new_buckets, 0);
        size_t n = items_;
        items_ = 0;
        growth_left_ = cap_of(new_buckets);
        for (size_t i = 0; i < ob; i++)
            if (!(octrl[i] & 0x80)) insert_no_grow(ohash[i], okey[i]);
        (void)n;
    }
};

}  // namespace orc
