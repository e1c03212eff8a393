// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// Literal emulation of tsl::robin_map 1.4.0 (cpp/kiss_icp/3rdparty/tsl_robin/tsl_robin.cmake:24;
// not vendored, not installed — restated from the upstream algorithm; PARITY UNPINNED) for
// the one property of it that is result-affecting in the reference: ITERATION ORDER.
//   * VoxelDownsample emits points in grid iteration order        (core/VoxelUtils.cpp:9-19)
//   * Pointcloud() concatenates voxels in map_ iteration order    (core/VoxelHashMap.cpp:75)
//   * RemovePointsFarFromLocation erases while iterating          (core/VoxelHashMap.cpp:123-131)
//
// Upstream semantics restated: power_of_two_growth_policy<2>, default bucket count 0,
// max_load_factor 0.5, min_load_factor 0, linear probing with robin-hood swap when
// (probe distance > resident distance), DIST_FROM_IDEAL_BUCKET_LIMIT = 8192, grow when
// size() >= load_threshold before an insert, rehash re-inserts in old bucket order,
// erase = clear + backward shift, iteration = bucket array order.
#pragma once

#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace oracle {

struct Voxel {
    int32_t x, y, z;
    bool operator==(const Voxel &o) const { return x == o.x && y == o.y && z == o.z; }
};

// std::hash<kiss_icp::Voxel>  (core/VoxelUtils.hpp:45-51): uint32 wrap-around arithmetic.
inline uint32_t voxel_hash(const Voxel &v) {
    return (static_cast<uint32_t>(v.x) * 73856093u) ^ (static_cast<uint32_t>(v.y) * 19349669u) ^
           (static_cast<uint32_t>(v.z) * 83492791u);
}

template <class V>
class RobinMap {
public:
    struct Bucket {
        int32_t dist = -1;  // -1 = empty (upstream: int16 dist_from_ideal_bucket)
        Voxel key{0, 0, 0};
        V value{};
        bool empty() const { return dist < 0; }
    };
    static constexpr int32_t kDistLimit = 8192;

    size_t size() const { return nb_elements_; }
    bool is_empty() const { return nb_elements_ == 0; }
    size_t bucket_count() const { return buckets_.size(); }
    const std::vector<Bucket> &buckets() const { return buckets_; }
    std::vector<Bucket> &buckets() { return buckets_; }

    // robin_hash::clear(): min_load_factor == 0 -> clear buckets in place, keep bucket_count
    void clear() {
        for (auto &b : buckets_) {
            b.dist = -1;
            b.value = V{};
        }
        nb_elements_ = 0;
        grow_on_next_insert_ = false;
    }

    // reserve(count) -> rehash(ceil(float(count)/max_load_factor))
    void reserve(size_t count) {
        size_t c = static_cast<size_t>(std::ceil(static_cast<float>(count) / max_load_factor_));
        const size_t min_c = static_cast<size_t>(std::ceil(static_cast<float>(size()) / max_load_factor_));
        if (c < min_c) c = min_c;
        rehash_impl(c);
    }

    // returns pointer to value or nullptr
    V *find(const Voxel &key) {
        if (buckets_.empty()) return nullptr;
        size_t ib = voxel_hash(key) & mask_;
        int32_t dist = 0;
        while (dist <= buckets_[ib].dist) {
            if (buckets_[ib].key == key) return &buckets_[ib].value;
            ib = (ib + 1) & mask_;
            ++dist;
        }
        return nullptr;
    }
    const V *find(const Voxel &key) const { return const_cast<RobinMap *>(this)->find(key); }
    // bucket index of `key` or bucket_count() (= end()): erase(find(key)) in the fuzz tests
    size_t find_index(const Voxel &key) const {
        if (buckets_.empty()) return 0;
        size_t ib = voxel_hash(key) & mask_;
        int32_t dist = 0;
        while (dist <= buckets_[ib].dist) {
            if (buckets_[ib].key == key) return ib;
            ib = (ib + 1) & mask_;
            ++dist;
        }
        return buckets_.size();
    }
    bool contains(const Voxel &key) const { return find(key) != nullptr; }

    // insert_impl: returns false if the key was already present.
    bool insert(const Voxel &key, V &&value) {
        const uint32_t hash = voxel_hash(key);
        size_t ib = buckets_.empty() ? 0 : (hash & mask_);
        int32_t dist = 0;
        if (!buckets_.empty()) {
            while (dist <= buckets_[ib].dist) {
                if (buckets_[ib].key == key) return false;
                ib = (ib + 1) & mask_;
                ++dist;
            }
        }
        while (rehash_on_extreme_load(dist)) {
            ib = hash & mask_;
            dist = 0;
            while (dist <= buckets_[ib].dist) {
                ib = (ib + 1) & mask_;
                ++dist;
            }
        }
        if (buckets_[ib].empty()) {
            buckets_[ib].dist = dist;
            buckets_[ib].key = key;
            buckets_[ib].value = std::move(value);
        } else {
            insert_value(ib, dist, key, std::move(value));
        }
        ++nb_elements_;
        return true;
    }

    // erase(iterator): clear + backward shift. Returns the bucket index iteration continues
    // from (upstream: `if (pos.m_bucket->empty()) ++pos; return pos;`), or bucket_count() = end.
    size_t erase_at(size_t ib) {
        buckets_[ib].dist = -1;
        buckets_[ib].value = V{};
        --nb_elements_;
        size_t prev = ib;
        size_t cur = (ib + 1) & mask_;
        while (buckets_[cur].dist > 0) {
            buckets_[prev].dist = buckets_[cur].dist - 1;
            buckets_[prev].key = buckets_[cur].key;
            buckets_[prev].value = std::move(buckets_[cur].value);
            buckets_[cur].dist = -1;
            buckets_[cur].value = V{};
            prev = cur;
            cur = (cur + 1) & mask_;
        }
        if (buckets_[ib].empty()) return next_occupied(ib + 1);
        return ib;
    }

    // first occupied bucket index >= from, or bucket_count() (= end())
    size_t next_occupied(size_t from) const {
        size_t i = from;
        while (i < buckets_.size() && buckets_[i].empty()) ++i;
        return i;
    }

private:
    bool rehash_on_extreme_load(int32_t curr_dist) {
        if (grow_on_next_insert_ || curr_dist > kDistLimit || size() >= load_threshold_) {
            // power_of_two_growth_policy::next_bucket_count(): (mask + 1) * 2, with mask = 0
            // for an empty table -> 2.
            rehash_impl((mask_ + 1) * 2);
            grow_on_next_insert_ = false;
            return true;
        }
        return false;  // min_load_factor == 0: the shrink branch is never taken
    }

    void insert_value(size_t ib, int32_t dist, Voxel key, V &&value) {
        std::swap(dist, buckets_[ib].dist);
        std::swap(key, buckets_[ib].key);
        std::swap(value, buckets_[ib].value);
        ib = (ib + 1) & mask_;
        ++dist;
        while (!buckets_[ib].empty()) {
            if (dist > buckets_[ib].dist) {
                if (dist > kDistLimit) grow_on_next_insert_ = true;
                std::swap(dist, buckets_[ib].dist);
                std::swap(key, buckets_[ib].key);
                std::swap(value, buckets_[ib].value);
            }
            ib = (ib + 1) & mask_;
            ++dist;
        }
        buckets_[ib].dist = dist;
        buckets_[ib].key = key;
        buckets_[ib].value = std::move(value);
    }

    static size_t round_up_pow2(size_t v) {
        if (v == 0) return 0;
        size_t p = 1;
        while (p < v) p <<= 1;
        return p;
    }

    void rehash_impl(size_t count) {
        const size_t nb = round_up_pow2(count);
        std::vector<Bucket> old;
        old.swap(buckets_);
        buckets_.assign(nb, Bucket{});
        mask_ = nb ? nb - 1 : 0;
        load_threshold_ = static_cast<size_t>(static_cast<float>(nb) * max_load_factor_);
        // insert_value_on_rehash in old bucket order
        for (auto &b : old) {
            if (b.empty()) continue;
            size_t ib = voxel_hash(b.key) & mask_;
            int32_t dist = 0;
            Voxel key = b.key;
            V value = std::move(b.value);
            while (true) {
                if (dist > buckets_[ib].dist) {
                    if (buckets_[ib].empty()) {
                        buckets_[ib].dist = dist;
                        buckets_[ib].key = key;
                        buckets_[ib].value = std::move(value);
                        break;
                    }
                    std::swap(dist, buckets_[ib].dist);
                    std::swap(key, buckets_[ib].key);
                    std::swap(value, buckets_[ib].value);
                }
                ++dist;
                ib = (ib + 1) & mask_;
            }
        }
    }

    std::vector<Bucket> buckets_;
    size_t mask_ = 0;
    size_t nb_elements_ = 0;
    size_t load_threshold_ = 0;
    float max_load_factor_ = 0.5f;
    bool grow_on_next_insert_ = false;
};

}  // namespace oracle
