// Test scaffolding (see oracle/shim/Eigen/Core): boost::unordered_map for integral keys with the NODE ORDER of
// boost 1.67-1.79 (the reference pins boost 1.75 through vcpkg), because the reference's output depends on the
// iteration order of its two hash maps:
//   * prime bucket policy, identity hash: bucket = key % bucket_count; default 17 buckets, allocated on first insert;
//     max load factor 1, growth to next_prime(max(size+1, size + size/2) + 1);
//   * all nodes on one singly linked list; a bucket stores the link PRECEDING its first node; inserting into an empty
//     bucket puts the node at the head of the whole list, otherwise right after the bucket's predecessor link;
//   * rehash walks the list from the head: a node whose new bucket is empty stays in place, otherwise it is moved
//     right after that bucket's predecessor link;
//   * erase unlinks one node and never rehashes; clear keeps the bucket count.
// The iterator type names under boost::unordered::{iterator_detail,detail} exist because the reference spells them
// out (ref: ASMC_SRC/SRC/HASHING/ExtendHash.hpp:40-43).
#pragma once
#include <algorithm>
#include <cstddef>
#include <utility>
#include <vector>

namespace boost
{
namespace unordered
{
namespace detail
{
struct link {
  link* next = nullptr;
};
template <class VT> struct ptr_node : link {
  VT value;
  std::size_t bucket;
  ptr_node(const VT& v, const std::size_t b) : value(v), bucket(b) {}
};
}  // namespace detail
namespace iterator_detail
{
template <class Node> struct iterator {
  Node* n = nullptr;
  iterator() = default;
  explicit iterator(Node* p) : n(p) {}
  auto& operator*() const { return n->value; }
  auto* operator->() const { return &n->value; }
  iterator& operator++()
  {
    n = static_cast<Node*>(n->next);
    return *this;
  }
  iterator operator++(int)
  {
    iterator old = *this;
    n = static_cast<Node*>(n->next);
    return old;
  }
  bool operator==(const iterator& o) const { return n == o.n; }
  bool operator!=(const iterator& o) const { return n != o.n; }
};
}  // namespace iterator_detail
}  // namespace unordered

template <class K, class V> class unordered_map
{
public:
  using value_type = std::pair<const K, V>;
  using node = unordered::detail::ptr_node<value_type>;
  using link = unordered::detail::link;
  using iterator = unordered::iterator_detail::iterator<node>;

private:
  link mHead;                     // sentinel: mHead.next = first node
  std::vector<link*> mBuckets;    // predecessor link of each bucket's first node, nullptr = empty bucket
  std::size_t mBucketCount = 17, mSize = 0, mMaxLoad = 0;

  static std::size_t nextPrime(const std::size_t n)
  {
    static const std::size_t primes[] = {
        17ul,       29ul,       37ul,        53ul,        67ul,        79ul,        97ul,        131ul,        193ul,       257ul,
        389ul,      521ul,      769ul,       1031ul,      1543ul,      2053ul,      3079ul,      6151ul,       12289ul,     24593ul,
        49157ul,    98317ul,    196613ul,    393241ul,    786433ul,    1572869ul,   3145739ul,   6291469ul,    12582917ul,  25165843ul,
        50331653ul, 100663319ul, 201326611ul, 402653189ul, 805306457ul, 1610612741ul, 3221225473ul, 4294967291ul};
    for (const std::size_t p : primes) {
      if (p >= n) {
        return p;
      }
    }
    return 4294967291ul;
  }
  void createBuckets(const std::size_t count)
  {
    mBucketCount = count;
    mBuckets.assign(count, nullptr);
    mMaxLoad = count;
  }
  void rehash(const std::size_t count)
  {
    createBuckets(count);
    link* prev = &mHead;
    while (prev->next) {
      node* n = static_cast<node*>(prev->next);
      const std::size_t b = static_cast<std::size_t>(n->value.first) % mBucketCount;
      n->bucket = b;
      if (!mBuckets[b]) {
        mBuckets[b] = prev;
        prev = n;
      } else {
        prev->next = n->next;
        n->next = mBuckets[b]->next;
        mBuckets[b]->next = n;
      }
    }
  }

public:
  unordered_map() = default;
  unordered_map(const unordered_map&) = delete;
  unordered_map& operator=(const unordered_map&) = delete;
  ~unordered_map() { freeNodes(); }

  std::size_t size() const { return mSize; }
  bool empty() const { return mSize == 0; }
  std::size_t bucket_count() const { return mBucketCount; }
  iterator begin() { return iterator(static_cast<node*>(mHead.next)); }
  iterator end() { return iterator(nullptr); }

  iterator find(const K& key)
  {
    if (!mBuckets.empty()) {
      const std::size_t b = static_cast<std::size_t>(key) % mBucketCount;
      if (mBuckets[b]) {
        for (node* n = static_cast<node*>(mBuckets[b]->next); n && n->bucket == b; n = static_cast<node*>(n->next)) {
          if (n->value.first == key) {
            return iterator(n);
          }
        }
      }
    }
    return end();
  }

  std::pair<iterator, bool> insert(const std::pair<K, V>& kv)
  {
    iterator found = find(kv.first);
    if (found != end()) {
      return {found, false};
    }
    if (mBuckets.empty()) {
      createBuckets(std::max(mBucketCount, nextPrime(mSize + 1 + 1)));
    } else if (mSize + 1 > mMaxLoad) {
      const std::size_t want = nextPrime(std::max(mSize + 1, mSize + (mSize >> 1)) + 1);
      if (want != mBucketCount) {
        rehash(want);
      }
    }
    const std::size_t b = static_cast<std::size_t>(kv.first) % mBucketCount;
    node* n = new node(value_type(kv.first, kv.second), b);
    if (!mBuckets[b]) {
      if (mHead.next) {
        mBuckets[static_cast<node*>(mHead.next)->bucket] = n;
      }
      mBuckets[b] = &mHead;
      n->next = mHead.next;
      mHead.next = n;
    } else {
      n->next = mBuckets[b]->next;
      mBuckets[b]->next = n;
    }
    ++mSize;
    return {iterator(n), true};
  }

  iterator erase(iterator it)
  {
    node* n = it.n;
    const std::size_t b = n->bucket;
    link* prev = mBuckets[b];
    while (prev->next != n) {
      prev = prev->next;
    }
    link* after = n->next;
    prev->next = after;
    --mSize;
    bool sameBucketFollows = false;
    if (after) {
      const std::size_t b2 = static_cast<node*>(after)->bucket;
      if (b2 == b) {
        sameBucketFollows = true;
      } else {
        mBuckets[b2] = prev;
      }
    }
    if (!sameBucketFollows && mBuckets[b] == prev) {
      mBuckets[b] = nullptr;
    }
    delete n;
    return iterator(static_cast<node*>(after));
  }

  void clear()
  {
    if (!mSize) {
      return;
    }
    freeNodes();
    std::fill(mBuckets.begin(), mBuckets.end(), nullptr);
  }

private:
  void freeNodes()
  {
    link* p = mHead.next;
    while (p) {
      link* nx = p->next;
      delete static_cast<node*>(p);
      p = nx;
    }
    mHead.next = nullptr;
    mSize = 0;
  }
};
}  // namespace boost
