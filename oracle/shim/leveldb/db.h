// Reference-harness stand-in for LevelDB 1.x: an ordered in-memory map with
// bytewise key order, persisted to "<path>/data.bin" on every Write so that the
// reference's --diff_init and --diff_repeat runs can be separate processes.
// TEST INFRASTRUCTURE ONLY: used to build the unmodified reference into oracle/_ref.
// A directory is created at the DB path because the reference walks it
// (leveldb_store.hpp:143-150).
#pragma once
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <string>
#include <utility>
#include <vector>
namespace leveldb {
class Slice {
 public:
  Slice() : d_(""), n_(0) {}
  Slice(const char* d, size_t n) : d_(d), n_(n) {}
  Slice(const std::string& s) : d_(s.data()), n_(s.size()) {}
  Slice(const char* s) : d_(s), n_(std::strlen(s)) {}
  const char* data() const { return d_; }
  size_t size() const { return n_; }
  std::string ToString() const { return std::string(d_, n_); }
  int compare(const Slice& b) const {
    size_t m = n_ < b.n_ ? n_ : b.n_;
    int r = std::memcmp(d_, b.d_, m);
    if (r == 0) { if (n_ < b.n_) r = -1; else if (n_ > b.n_) r = 1; }
    return r;
  }
 private:
  const char* d_;
  size_t n_;
};
class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& m) : ok_(false), msg_(m) {}
  bool ok() const { return ok_; }
  std::string ToString() const { return ok_ ? "OK" : msg_; }
 private:
  bool ok_;
  std::string msg_;
};
struct Options {
  bool create_if_missing = false;
  bool error_if_exists = false;
  size_t write_buffer_size = 4 << 20;
};
struct ReadOptions {};
struct WriteOptions { bool sync = false; };
class WriteBatch {
 public:
  void Put(const Slice& k, const Slice& v) { ops_.emplace_back(k.ToString(), v.ToString()); }
  std::vector<std::pair<std::string, std::string> > ops_;
};
class Iterator {
 public:
  explicit Iterator(const std::map<std::string, std::string>* m) : m_(m), it_(m->end()), valid_(false) {}
  void Seek(const Slice& k) { it_ = m_->lower_bound(k.ToString()); valid_ = it_ != m_->end(); }
  void SeekToFirst() { it_ = m_->begin(); valid_ = it_ != m_->end(); }
  void SeekToLast() { if (m_->empty()) { valid_ = false; return; } it_ = m_->end(); --it_; valid_ = true; }
  void Next() { if (!valid_) return; ++it_; valid_ = it_ != m_->end(); }
  void Prev() { if (!valid_) return; if (it_ == m_->begin()) { valid_ = false; return; } --it_; }
  bool Valid() const { return valid_; }
  Slice key() const { return Slice(it_->first); }
  Slice value() const { return Slice(it_->second); }
 private:
  const std::map<std::string, std::string>* m_;
  std::map<std::string, std::string>::const_iterator it_;
  bool valid_;
};
class DB {
 public:
  static Status Open(const Options& o, const std::string& path, DB** out) {
    namespace fs = std::filesystem;
    bool exists = fs::exists(path + "/data.bin");
    if (exists && o.error_if_exists) return Status("Invalid argument: " + path + ": exists (error_if_exists is true)");
    if (!exists && !o.create_if_missing) return Status("Invalid argument: " + path + ": does not exist (create_if_missing is false)");
    std::error_code ec;
    fs::create_directories(path, ec);
    DB* db = new DB(path);
    if (exists) db->load(); else db->flush();
    *out = db;
    return Status();
  }
  Status Write(const WriteOptions&, WriteBatch* b) {
    for (auto& kv : b->ops_) m_[kv.first] = kv.second;
    b->ops_.clear();
    flush();
    return Status();
  }
  Status Put(const WriteOptions& w, const Slice& k, const Slice& v) {
    m_[k.ToString()] = v.ToString();
    (void)w; flush();
    return Status();
  }
  Status Get(const ReadOptions&, const Slice& k, std::string* v) {
    auto it = m_.find(k.ToString());
    if (it == m_.end()) return Status("NotFound");
    *v = it->second;
    return Status();
  }
  Iterator* NewIterator(const ReadOptions&) { return new Iterator(&m_); }
 private:
  explicit DB(const std::string& p) : path_(p) {}
  void flush() {
    std::ofstream f(path_ + "/data.bin", std::ios::binary | std::ios::trunc);
    for (auto& kv : m_) {
      std::uint64_t a = kv.first.size(), b = kv.second.size();
      f.write(reinterpret_cast<const char*>(&a), 8); f.write(kv.first.data(), a);
      f.write(reinterpret_cast<const char*>(&b), 8); f.write(kv.second.data(), b);
    }
  }
  void load() {
    std::ifstream f(path_ + "/data.bin", std::ios::binary);
    for (;;) {
      std::uint64_t a = 0, b = 0;
      if (!f.read(reinterpret_cast<char*>(&a), 8)) break;
      std::string k(a, '\0'); f.read(&k[0], a);
      f.read(reinterpret_cast<char*>(&b), 8);
      std::string v(b, '\0'); f.read(&v[0], b);
      m_[k] = v;
    }
  }
  std::string path_;
  std::map<std::string, std::string> m_;
};
}  // namespace leveldb
