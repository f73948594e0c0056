// ORACLE TEST TOOLING ONLY (never linked into the product library).
// Minimal C++17-standard-library stand-in for the subset of abseil-cpp 20211102.0 that the
// reference's open_spiel core + games/coup.cc use. abseil is un-vendored in /root/reference
// (open_spiel/scripts/install.sh:114-116 clones it at install time) and there is no network here.
// No game arithmetic lives in abseil: it only supplies containers/strings/Span/optional.
#ifndef ABSL_SHIM_ALL_H_
#define ABSL_SHIM_ALL_H_
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iomanip>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <optional>
#include <random>
#include <sstream>
#include <string>
#include <string_view>
#include <type_traits>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#define ABSL_DEPRECATED(msg)
#define ABSL_MUST_USE_RESULT
#define ABSL_GUARDED_BY(x)
#define ABSL_ATTRIBUTE_UNUSED __attribute__((unused))

namespace absl {

// ---- optional
template <class T> using optional = std::optional<T>;
using nullopt_t = std::nullopt_t;
inline constexpr std::nullopt_t nullopt = std::nullopt;

// ---- string_view
using string_view = std::string_view;

// ---- Span
template <class T>
class Span {
 public:
  using value_type = std::remove_cv_t<T>;
  using iterator = T*;
  using const_iterator = const T*;
  using size_type = size_t;
  Span() : p_(nullptr), n_(0) {}
  Span(T* p, size_t n) : p_(p), n_(n) {}
  template <class V, class = std::enable_if_t<
                         std::is_same_v<typename V::value_type, value_type> &&
                         !std::is_same_v<std::decay_t<V>, Span<T>>>>
  Span(V& v) : p_(v.data()), n_(v.size()) {}
  template <class V, class = std::enable_if_t<
                         std::is_const_v<T> &&
                         std::is_same_v<typename V::value_type, value_type>>,
            class = void>
  Span(const V& v) : p_(v.data()), n_(v.size()) {}
  template <class U = T, class = std::enable_if_t<std::is_const_v<U>>>
  Span(std::initializer_list<value_type> il) : p_(il.begin()), n_(il.size()) {}
  T* data() const { return p_; }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  T& operator[](size_t i) const { return p_[i]; }
  T& at(size_t i) const { return p_[i]; }
  T* begin() const { return p_; }
  T* end() const { return p_ + n_; }
  T& front() const { return p_[0]; }
  T& back() const { return p_[n_ - 1]; }
  Span subspan(size_t pos, size_t len = static_cast<size_t>(-1)) const {
    if (pos > n_) std::abort();
    return Span(p_ + pos, std::min(len, n_ - pos));
  }

 private:
  T* p_;
  size_t n_;
};
template <class V>
auto MakeSpan(V& v) { return Span<std::remove_reference_t<decltype(*v.data())>>(v.data(), v.size()); }
template <class T>
Span<T> MakeSpan(T* p, size_t n) { return Span<T>(p, n); }
template <class V>
auto MakeConstSpan(const V& v) { return Span<const typename V::value_type>(v.data(), v.size()); }

// ---- StrCat / StrAppend
namespace shim_internal {
template <class T>
void Put(std::ostream& os, const T& v) {
  if constexpr (std::is_enum_v<T>) {
    os << static_cast<long long>(static_cast<std::underlying_type_t<T>>(v));
  } else if constexpr (std::is_same_v<T, bool>) {
    os << (v ? 1 : 0);
  } else {
    os << v;
  }
}
}  // namespace shim_internal
template <class... A>
std::string StrCat(const A&... a) {
  std::ostringstream os;
  (shim_internal::Put(os, a), ...);
  return os.str();
}
template <class... A>
void StrAppend(std::string* s, const A&... a) { s->append(StrCat(a...)); }

// ---- StrJoin
struct AlphaNumFormatterImpl {
  template <class T> void operator()(std::string* out, const T& v) const { StrAppend(out, v); }
};
inline AlphaNumFormatterImpl AlphaNumFormatter() { return {}; }
template <class F1, class F2>
struct PairFormatterImpl {
  F1 f1; std::string sep; F2 f2;
  template <class P> void operator()(std::string* out, const P& p) const {
    f1(out, p.first); out->append(sep); f2(out, p.second);
  }
};
template <class F1, class F2>
PairFormatterImpl<F1, F2> PairFormatter(F1 f1, string_view sep, F2 f2) { return {f1, std::string(sep), f2}; }
inline auto PairFormatter(string_view sep) { return PairFormatter(AlphaNumFormatter(), sep, AlphaNumFormatter()); }
template <class C, class F>
std::string StrJoin(const C& c, string_view sep, F f) {
  std::string out; bool first = true;
  for (const auto& v : c) { if (!first) out.append(sep); first = false; f(&out, v); }
  return out;
}
template <class C>
std::string StrJoin(const C& c, string_view sep) { return StrJoin(c, sep, AlphaNumFormatter()); }
template <class T>
std::string StrJoin(std::initializer_list<T> c, string_view sep) { return StrJoin<std::initializer_list<T>>(c, sep, AlphaNumFormatter()); }

// ---- StrSplit
struct MaxSplitsImpl { std::string delim; int limit; };
inline MaxSplitsImpl MaxSplits(string_view d, int limit) { return {std::string(d), limit}; }
inline MaxSplitsImpl MaxSplits(char d, int limit) { return {std::string(1, d), limit}; }
class Splitter {
 public:
  Splitter(string_view text, const std::string& delim, int limit) {
    size_t pos = 0; int splits = 0;
    while (true) {
      size_t f = (limit >= 0 && splits >= limit) ? std::string::npos : text.find(delim, pos);
      if (delim.empty()) f = std::string::npos;
      if (f == std::string::npos) { parts_.emplace_back(text.substr(pos)); break; }
      parts_.emplace_back(text.substr(pos, f - pos));
      pos = f + delim.size(); ++splits;
    }
  }
  operator std::vector<std::string>() const { return std::vector<std::string>(parts_.begin(), parts_.end()); }
  operator std::vector<string_view>() const { return parts_; }
  template <class A, class B>
  operator std::pair<A, B>() const {
    return {A(parts_.size() > 0 ? parts_[0] : string_view()), B(parts_.size() > 1 ? parts_[1] : string_view())};
  }
  auto begin() const { return parts_.begin(); }
  auto end() const { return parts_.end(); }
 private:
  std::vector<string_view> parts_;  // views into caller-owned text
};
inline Splitter StrSplit(string_view t, string_view d) { return Splitter(t, std::string(d), -1); }
inline Splitter StrSplit(string_view t, char d) { return Splitter(t, std::string(1, d), -1); }
inline Splitter StrSplit(string_view t, const MaxSplitsImpl& m) { return Splitter(t, m.delim, m.limit); }

// ---- StrFormat (tiny printf subset: %s %d %i %f %g %a with width/precision)
namespace shim_internal {
inline void FormatRest(std::ostringstream& os, const char* f) {
  while (*f) { if (f[0] == '%' && f[1] == '%') { os << '%'; f += 2; } else os << *f++; }
}
template <class T, class... R>
void FormatRest(std::ostringstream& os, const char* f, const T& v, const R&... r) {
  while (*f) {
    if (f[0] == '%' && f[1] == '%') { os << '%'; f += 2; continue; }
    if (*f != '%') { os << *f++; continue; }
    std::string spec = "%"; ++f;
    while (*f && !std::strchr("sdifgeaxuc", *f)) spec += *f++;
    char conv = *f++;
    char buf[512];
    if constexpr (std::is_floating_point_v<T>) {
      spec += (conv == 's' || conv == 'd' || conv == 'i') ? 'g' : conv;
      std::snprintf(buf, sizeof buf, spec.c_str(), static_cast<double>(v)); os << buf;
    } else if constexpr (std::is_integral_v<T> || std::is_enum_v<T>) {
      if (conv == 'f' || conv == 'g' || conv == 'e') { spec += conv; std::snprintf(buf, sizeof buf, spec.c_str(), static_cast<double>(v)); }
      else if (conv == 'c') { spec += 'c'; std::snprintf(buf, sizeof buf, spec.c_str(), static_cast<int>(v)); }
      else { spec += "lld"; std::snprintf(buf, sizeof buf, spec.c_str(), static_cast<long long>(v)); }
      os << buf;
    } else {
      std::ostringstream t; t << v; os << t.str();
    }
    FormatRest(os, f, r...);
    return;
  }
}
}  // namespace shim_internal
template <class... A>
std::string StrFormat(const char* f, const A&... a) {
  std::ostringstream os; shim_internal::FormatRest(os, f, a...); return os.str();
}
template <class... A>
std::string StreamFormat(const char* f, const A&... a) { return StrFormat(f, a...); }

// ---- match / numbers / replace / ascii / charconv
inline bool StrContains(string_view h, string_view n) { return h.find(n) != string_view::npos; }
inline bool StrContains(string_view h, char n) { return h.find(n) != string_view::npos; }
inline bool StartsWith(string_view t, string_view p) { return t.substr(0, p.size()) == p; }
inline bool EndsWith(string_view t, string_view s) { return t.size() >= s.size() && t.substr(t.size() - s.size()) == s; }
template <class I>
bool SimpleAtoi(string_view s, I* out) {
  std::string t(s); char* e = nullptr; errno = 0;
  long long v = std::strtoll(t.c_str(), &e, 10);
  if (t.empty() || *e != '\0' || errno) return false;
  *out = static_cast<I>(v); return true;
}
inline bool SimpleAtod(string_view s, double* out) {
  std::string t(s); char* e = nullptr; errno = 0;
  double v = std::strtod(t.c_str(), &e);
  if (t.empty() || *e != '\0') return false;
  *out = v; return true;
}
inline std::string StrReplaceAll(string_view s, std::initializer_list<std::pair<string_view, string_view>> reps) {
  std::string out(s);
  for (auto& [from, to] : reps) {
    if (from.empty()) continue;
    size_t pos = 0;
    while ((pos = out.find(from, pos)) != std::string::npos) { out.replace(pos, from.size(), to); pos += to.size(); }
  }
  return out;
}
inline string_view StripAsciiWhitespace(string_view s) {
  size_t b = 0, e = s.size();
  while (b < e && std::isspace(static_cast<unsigned char>(s[b]))) ++b;
  while (e > b && std::isspace(static_cast<unsigned char>(s[e - 1]))) --e;
  return s.substr(b, e - b);
}
inline bool ascii_isspace(unsigned char c) { return std::isspace(c); }
inline bool ascii_isdigit(unsigned char c) { return std::isdigit(c); }
inline std::string AsciiStrToLower(string_view s) { std::string o(s); for (auto& c : o) c = std::tolower(c); return o; }
using from_chars_result = std::from_chars_result;
inline from_chars_result from_chars(const char* b, const char* e, double& v) { return std::from_chars(b, e, v); }

// ---- algorithm/container
template <class C, class T> T c_accumulate(const C& c, T init) { return std::accumulate(c.begin(), c.end(), init); }
template <class C, class T, class Op> T c_accumulate(const C& c, T init, Op op) { return std::accumulate(c.begin(), c.end(), init, op); }
template <class C, class F> void c_for_each(C&& c, F f) { std::for_each(c.begin(), c.end(), f); }
template <class C, class T> auto c_find(C& c, const T& v) { return std::find(c.begin(), c.end(), v); }
template <class C, class P> auto c_find_if(C& c, P p) { return std::find_if(c.begin(), c.end(), p); }
template <class C, class P> bool c_all_of(const C& c, P p) { return std::all_of(c.begin(), c.end(), p); }
template <class C, class P> bool c_any_of(const C& c, P p) { return std::any_of(c.begin(), c.end(), p); }
template <class C, class T> void c_fill(C&& c, const T& v) { std::fill(c.begin(), c.end(), v); }
template <class C, class T> void c_iota(C& c, T v) { std::iota(c.begin(), c.end(), v); }
template <class C> void c_sort(C& c) { std::sort(c.begin(), c.end()); }
template <class C, class Cmp> void c_sort(C& c, Cmp cmp) { std::sort(c.begin(), c.end(), cmp); }
template <class C, class T> auto c_count(const C& c, const T& v) { return std::count(c.begin(), c.end(), v); }
template <class C> auto c_max_element(C& c) { return std::max_element(c.begin(), c.end()); }
template <class C> auto c_min_element(C& c) { return std::min_element(c.begin(), c.end()); }
template <class C, class T> bool c_linear_search(const C& c, const T& v) { return std::find(c.begin(), c.end(), v) != c.end(); }

// ---- memory
template <class T, class... A> std::unique_ptr<T> make_unique(A&&... a) { return std::make_unique<T>(std::forward<A>(a)...); }

// ---- containers
template <class T, size_t N> using InlinedVector = std::vector<T>;
template <class K, class H = std::hash<K>> using flat_hash_set = std::unordered_set<K, H>;
template <class K, class V, class H = std::hash<K>> using flat_hash_map = std::unordered_map<K, V, H>;

// ---- random
class BitGenRef {
 public:
  using result_type = uint64_t;
  template <class G, class = std::enable_if_t<!std::is_same_v<std::decay_t<G>, BitGenRef>>>
  BitGenRef(G& g) : fn_([&g]() -> uint64_t {
      if constexpr (sizeof(typename G::result_type) >= 8 && G::max() == ~0ull) return g();
      else return (static_cast<uint64_t>(g()) << 32) ^ static_cast<uint64_t>(g());
    }) {}
  static constexpr uint64_t min() { return 0; }
  static constexpr uint64_t max() { return ~0ull; }
  uint64_t operator()() { return fn_(); }
 private:
  std::function<uint64_t()> fn_;
};
template <class G>
double Uniform(G&& g, double lo, double hi) { return std::uniform_real_distribution<double>(lo, hi)(g); }
template <class G>
int Uniform(G&& g, int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi - 1)(g); }
template <class T = double> using uniform_real_distribution = std::uniform_real_distribution<T>;
template <class T = int> using uniform_int_distribution = std::uniform_int_distribution<T>;
template <class T = int> using discrete_distribution = std::discrete_distribution<T>;
using BitGen = std::mt19937_64;

// ---- time
struct Duration { int64_t ns; };
struct Time { int64_t ns; };
inline Time Now() { return {std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count()}; }
inline Time UnixEpoch() { return {0}; }
inline Duration operator-(Time a, Time b) { return {a.ns - b.ns}; }
inline int64_t ToInt64Nanoseconds(Duration d) { return d.ns; }
inline int64_t ToInt64Milliseconds(Duration d) { return d.ns / 1000000; }
inline double ToDoubleSeconds(Duration d) { return d.ns * 1e-9; }

// ---- mutex
class Mutex { public: void Lock() { m_.lock(); } void Unlock() { m_.unlock(); } private: std::mutex m_; };
class MutexLock { public: explicit MutexLock(Mutex* m) : m_(m) { m_->Lock(); } ~MutexLock() { m_->Unlock(); } private: Mutex* m_; };

}  // namespace absl
#endif
