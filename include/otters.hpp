// otters.hpp — header-only C++17 host facade over the C ABI (include/otters_b200.h).
//
// Mirrors the reference's public Rust API (src/prelude.rs:7-23) name for name so that its tests read the
// same: VecStore / VecQueryPlan (src/vec.rs), MetaStore / MetaStoreBuilder / MetaQueryPlan (src/meta.rs),
// Column / DataType (src/col.rs, src/type_utils.rs:11-19), Expr / col() / lit() / compile (src/expr.rs).
// Errors are deferred to collect()/build() and surface as Result<T> carrying the reference's message strings.
// All scoring, filtering and selection happens in libotters_b200.so on the GPU; this header only builds plans.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "otters_b200.h"

namespace otters {

// ---- Result<T, String> -------------------------------------------------------------------------------------
template <typename T>
class Result {
    std::optional<T> v_;
    std::string e_;

  public:
    static Result Ok(T v) { Result r; r.v_ = std::move(v); return r; }
    static Result Err(std::string e) { Result r; r.e_ = std::move(e); return r; }
    bool is_ok() const { return v_.has_value(); }
    bool is_err() const { return !v_.has_value(); }
    T& unwrap() { if (!v_) throw std::runtime_error("called unwrap() on Err: " + e_); return *v_; }
    const std::string& unwrap_err() const { return e_; }
};

enum class Metric : int32_t { Cosine = 0, Euclidean = 1, DotProduct = 2 };          // src/vec.rs:11-16
enum class TakeType : int32_t { Min = 0, Max = 1 };                                 // src/vec.rs:18-22
enum class Cmp : int32_t { Lt = 0, Gt = 1, Lte = 2, Gte = 3, Eq = 4 };              // src/vec.rs:24-31
enum class CmpOp : int32_t { Eq = 0, Neq = 1, Lt = 2, Lte = 3, Gt = 4, Gte = 5 };   // src/expr.rs:83-91
enum class DataType : int32_t { Int32 = 0, Int64 = 1, Float32 = 2, Float64 = 3, String = 4, DateTime = 5 };

inline const char* dtype_name(DataType d) {
    static const char* n[] = {"Int32", "Int64", "Float32", "Float64", "String", "DateTime"};
    return n[(int)d];
}

struct SearchResult {  // src/vec.rs:33-53
    size_t index;
    float score;
};

inline TakeType infer_default_take_type(Metric m) { return m == Metric::Euclidean ? TakeType::Min : TakeType::Max; }

// ---- device context (one per process and device, created on first use) ---------------------------------------
inline otters_ctx* default_ctx() {
    static otters_ctx* ctx = [] {
        otters_ctx* c = nullptr;
        if (otters_ctx_create(0, nullptr, &c) != OTTERS_OK) throw std::runtime_error(otters_last_error());
        return c;
    }();
    return ctx;
}

// ---- datetime literals: RFC3339 | YYYY-MM-DD | YYYY-MM-DD HH:MM:SS -> epoch millis UTC (src/col.rs:506-529) ----
inline int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (int64_t)doe - 719468;
}
inline bool valid_date(int y, int mo, int d) {
    if (mo < 1 || mo > 12 || d < 1) return false;
    static const int dm[] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    int lim = dm[mo - 1] + ((mo == 2 && ((y % 4 == 0 && y % 100 != 0) || y % 400 == 0)) ? 1 : 0);
    return d <= lim;
}
inline std::optional<int64_t> parse_datetime_millis(const std::string& s) {
    int y, mo, d, h, mi, sec, n = 0;
    auto base = [&](int hh, int mm, int ss) { return (days_from_civil(y, mo, d) * 86400 + hh * 3600 + mm * 60 + ss) * 1000; };
    if (s.size() >= 20 && std::sscanf(s.c_str(), "%4d-%2d-%2d%*1[Tt ]%2d:%2d:%2d%n", &y, &mo, &d, &h, &mi, &sec, &n) == 6 && n == 19) {
        if (!valid_date(y, mo, d) || h > 23 || mi > 59 || sec > 60) return std::nullopt;
        size_t p = 19;
        int64_t frac_ms = 0;
        if (p < s.size() && s[p] == '.') {
            ++p;
            int digits = 0;
            int64_t v = 0;
            while (p < s.size() && s[p] >= '0' && s[p] <= '9') {
                if (digits < 3) { v = v * 10 + (s[p] - '0'); ++digits; }
                ++p;
            }
            if (digits == 0) return std::nullopt;
            while (digits++ < 3) v *= 10;
            frac_ms = v;
        }
        if (p >= s.size()) return std::nullopt;
        int64_t off = 0;
        if (s[p] == 'Z' || s[p] == 'z') {
            if (p + 1 != s.size()) return std::nullopt;
        } else if ((s[p] == '+' || s[p] == '-') && p + 6 == s.size() && s[p + 3] == ':') {
            int oh = (s[p + 1] - '0') * 10 + (s[p + 2] - '0'), om = (s[p + 4] - '0') * 10 + (s[p + 5] - '0');
            off = (s[p] == '+' ? 1 : -1) * (oh * 3600 + om * 60) * 1000LL;
        } else return std::nullopt;
        return base(h, mi, sec) + frac_ms - off;
    }
    if (s.size() == 10 && std::sscanf(s.c_str(), "%4d-%2d-%2d%n", &y, &mo, &d, &n) == 3 && n == 10)
        return valid_date(y, mo, d) ? std::optional<int64_t>(base(0, 0, 0)) : std::nullopt;
    if (s.size() == 19 && std::sscanf(s.c_str(), "%4d-%2d-%2d %2d:%2d:%2d%n", &y, &mo, &d, &h, &mi, &sec, &n) == 6 && n == 19)
        return (valid_date(y, mo, d) && h <= 23 && mi <= 59 && sec <= 59) ? std::optional<int64_t>(base(h, mi, sec)) : std::nullopt;
    return std::nullopt;
}

// ---- Column (src/col.rs) ---------------------------------------------------------------------------------------
class Column {
    std::string name_;
    DataType dtype_;
    std::vector<int32_t> i32_;
    std::vector<int64_t> i64_;  // Int64 and DateTime
    std::vector<float> f32_;
    std::vector<double> f64_;
    std::vector<std::string> str_;
    std::vector<uint8_t> nulls_;  // 1 = NULL (src/col.rs:26)

  public:
    Column(std::string name, DataType dt) : name_(std::move(name)), dtype_(dt) {}
    const std::string& name() const { return name_; }
    DataType dtype() const { return dtype_; }
    size_t len() const { return nulls_.size(); }
    bool is_empty() const { return nulls_.empty(); }
    const std::vector<uint8_t>& null_mask() const { return nulls_; }

    // NULL with the reference's sentinels (src/col.rs:238-326)
    void push_null() {
        nulls_.push_back(1);
        switch (dtype_) {
        case DataType::Int32: i32_.push_back(INT32_MIN); break;
        case DataType::Int64: case DataType::DateTime: i64_.push_back(INT64_MIN); break;
        case DataType::Float32: f32_.push_back(NAN); break;
        case DataType::Float64: f64_.push_back(NAN); break;
        case DataType::String: str_.emplace_back(); break;
        }
    }
    Result<bool> push(int32_t v) { return dtype_ == DataType::Int32 ? (i32_.push_back(v), nulls_.push_back(0), Result<bool>::Ok(true)) : mismatch(); }
    Result<bool> push(int64_t v) {
        if (dtype_ != DataType::Int64 && dtype_ != DataType::DateTime) return mismatch();
        i64_.push_back(v); nulls_.push_back(0); return Result<bool>::Ok(true);
    }
    Result<bool> push(float v) { return dtype_ == DataType::Float32 ? (f32_.push_back(v), nulls_.push_back(0), Result<bool>::Ok(true)) : mismatch(); }
    Result<bool> push(double v) { return dtype_ == DataType::Float64 ? (f64_.push_back(v), nulls_.push_back(0), Result<bool>::Ok(true)) : mismatch(); }
    Result<bool> push(const std::string& v) {
        if (dtype_ == DataType::String) { str_.push_back(v); nulls_.push_back(0); return Result<bool>::Ok(true); }
        if (dtype_ == DataType::DateTime) {  // strings are parsed as datetimes (src/col.rs:371-383)
            auto ms = parse_datetime_millis(v);
            if (!ms) return Result<bool>::Err("Parse error: Cannot parse '" + v + "' as datetime. Supported formats: ISO 8601, YYYY-MM-DD, YYYY-MM-DD HH:MM:SS");
            return push(*ms);
        }
        return mismatch();
    }
    Result<bool> push(const char* v) { return push(std::string(v)); }
    template <typename T>
    Result<bool> push(const std::optional<T>& v) { if (!v) { push_null(); return Result<bool>::Ok(true); } return push(*v); }
    template <typename T>
    Column from(const std::vector<T>& vals) && {  // Column::from (src/col.rs:392-401); throws on a type mismatch
        for (const auto& v : vals) { auto r = push(v); if (r.is_err()) throw std::runtime_error(r.unwrap_err()); }
        return std::move(*this);
    }

    // ABI view
    void fill(otters_column* c, std::vector<uint64_t>* null_words, std::vector<uint64_t>* offs, std::vector<uint8_t>* bytes) const {
        c->name = name_.c_str();
        c->dtype = (int32_t)dtype_;
        c->values = nullptr; c->null_words = nullptr; c->str_offsets = nullptr; c->str_bytes = nullptr;
        bool any = false;
        null_words->assign((len() + 63) / 64, 0);
        for (size_t i = 0; i < len(); ++i) if (nulls_[i]) { (*null_words)[i >> 6] |= 1ull << (i & 63); any = true; }
        if (any) c->null_words = null_words->data();
        switch (dtype_) {
        case DataType::Int32: c->values = i32_.data(); break;
        case DataType::Int64: case DataType::DateTime: c->values = i64_.data(); break;
        case DataType::Float32: c->values = f32_.data(); break;
        case DataType::Float64: c->values = f64_.data(); break;
        case DataType::String:
            offs->assign(1, 0);
            bytes->clear();
            for (const auto& s : str_) { bytes->insert(bytes->end(), s.begin(), s.end()); offs->push_back(bytes->size()); }
            if (bytes->empty()) bytes->push_back(0);
            c->str_offsets = offs->data();
            c->str_bytes = bytes->data();
            break;
        }
    }
    Column gather(const std::vector<size_t>& idx) const {  // src/meta.rs:723-821
        Column out(name_, dtype_);
        for (size_t i : idx) {
            out.nulls_.push_back(nulls_[i]);
            switch (dtype_) {
            case DataType::Int32: out.i32_.push_back(i32_[i]); break;
            case DataType::Int64: case DataType::DateTime: out.i64_.push_back(i64_[i]); break;
            case DataType::Float32: out.f32_.push_back(f32_[i]); break;
            case DataType::Float64: out.f64_.push_back(f64_[i]); break;
            case DataType::String: out.str_.push_back(str_[i]); break;
            }
        }
        return out;
    }
    const std::vector<int32_t>& i32_values() const { return i32_; }
    const std::vector<int64_t>& i64_values() const { return i64_; }
    const std::vector<double>& f64_values() const { return f64_; }
    const std::vector<std::string>& string_values() const { return str_; }

  private:
    Result<bool> mismatch() const { return Result<bool>::Err(std::string("Type mismatch: expected ") + dtype_name(dtype_) + ", got incompatible type"); }
};

// ---- expressions (src/expr.rs) -----------------------------------------------------------------------------------
using Literal = std::variant<int64_t, double, std::string>;
struct ColumnFilter {  // src/expr.rs:192-210
    std::string column;
    CmpOp cmp;
    Literal rhs;
    bool operator==(const ColumnFilter& o) const { return column == o.column && cmp == o.cmp && rhs == o.rhs; }
};
struct CompiledFilter { std::vector<std::vector<ColumnFilter>> clauses; };  // AND of OR-clauses (src/expr.rs:212-226)

class Expr {
  public:
    enum Kind { Col, Lit, CmpK, And, Or } kind;
    std::string name;
    Literal lit_v;
    CmpOp op = CmpOp::Eq;
    std::shared_ptr<Expr> a, b;

    static Expr column(std::string n) { Expr e; e.kind = Col; e.name = std::move(n); return e; }
    static Expr literal(Literal l) { Expr e; e.kind = Lit; e.lit_v = std::move(l); return e; }
    Expr cmp(CmpOp o, Literal l) const { Expr e; e.kind = CmpK; e.op = o; e.a = std::make_shared<Expr>(*this); e.b = std::make_shared<Expr>(literal(std::move(l))); return e; }
    template <typename T> static Literal to_lit(T v) {
        if constexpr (std::is_integral_v<T>) return (int64_t)v;
        else if constexpr (std::is_floating_point_v<T>) return (double)v;
        else return std::string(v);
    }
    template <typename T> Expr eq(T v) const { return cmp(CmpOp::Eq, to_lit(v)); }
    template <typename T> Expr neq(T v) const { return cmp(CmpOp::Neq, to_lit(v)); }
    template <typename T> Expr lt(T v) const { return cmp(CmpOp::Lt, to_lit(v)); }
    template <typename T> Expr lte(T v) const { return cmp(CmpOp::Lte, to_lit(v)); }
    template <typename T> Expr gt(T v) const { return cmp(CmpOp::Gt, to_lit(v)); }
    template <typename T> Expr gte(T v) const { return cmp(CmpOp::Gte, to_lit(v)); }
    Expr operator&(const Expr& o) const { Expr e; e.kind = And; e.a = std::make_shared<Expr>(*this); e.b = std::make_shared<Expr>(o); return e; }
    Expr operator|(const Expr& o) const { Expr e; e.kind = Or; e.a = std::make_shared<Expr>(*this); e.b = std::make_shared<Expr>(o); return e; }

    using Plan = std::vector<std::vector<ColumnFilter>>;
    Result<CompiledFilter> compile(const std::map<std::string, DataType>& schema) const {  // src/expr.rs:285-297
        std::string err;
        Plan p = lower(schema, &err);
        if (!err.empty()) return Result<CompiledFilter>::Err(err);
        CompiledFilter cf;
        for (auto& clause : p) {  // normalize_plan: drop (c == v) OR (c != v) (src/expr.rs:302-343)
            bool taut = false;
            for (auto& lf : clause)
                if (lf.cmp == CmpOp::Eq)
                    for (auto& x : clause)
                        if (x.cmp == CmpOp::Neq && x.column == lf.column && x.rhs == lf.rhs) taut = true;
            if (!taut) cf.clauses.push_back(clause);
        }
        return Result<CompiledFilter>::Ok(std::move(cf));
    }

  private:
    Plan lower(const std::map<std::string, DataType>& schema, std::string* err) const {  // src/expr.rs:355-372
        if (!err->empty()) return {};
        if (kind == And || kind == Or) {
            Plan l = a->lower(schema, err), r = b->lower(schema, err);
            if (!err->empty()) return {};
            if (l.empty()) return r;
            if (r.empty()) return l;
            if (kind == And) { l.insert(l.end(), r.begin(), r.end()); return l; }
            Plan out;  // or_distribute_clauses (src/expr.rs:494-511)
            for (auto& ca : l) for (auto& cb : r) { auto m = ca; m.insert(m.end(), cb.begin(), cb.end()); out.push_back(std::move(m)); }
            return out;
        }
        if (kind != CmpK) { *err = "Invalid expression (unexpected literal or column without comparator)"; return {}; }
        if (a->kind != Col || b->kind != Lit) { *err = "Invalid expression shape for comparison (expect column vs literal)"; return {}; }
        auto it = schema.find(a->name);
        if (it == schema.end()) { *err = "Unknown column '" + a->name + "'"; return {}; }
        const DataType dt = it->second;
        const Literal& l = b->lit_v;
        auto mismatch = [&](const char* got) { *err = "Type mismatch for column '" + a->name + "': expected " + dtype_name(dt) + ", got literal " + got; return Plan{}; };
        ColumnFilter f{a->name, op, l};
        switch (dt) {  // compile_cmp_leaf (src/expr.rs:385-466)
        case DataType::String:
            if (op != CmpOp::Eq && op != CmpOp::Neq) { *err = "Unsupported comparator for string column '" + a->name + "'"; return {}; }
            if (!std::holds_alternative<std::string>(l)) return mismatch("string");
            break;
        case DataType::Int32: case DataType::Int64:
            if (std::holds_alternative<double>(l)) return mismatch("float");
            if (std::holds_alternative<std::string>(l)) return mismatch("string");
            break;
        case DataType::DateTime: {
            if (!std::holds_alternative<std::string>(l)) return mismatch("datetime string");
            auto ms = parse_datetime_millis(std::get<std::string>(l));
            if (!ms) return mismatch("datetime string");
            f.rhs = *ms;
            break;
        }
        case DataType::Float32: case DataType::Float64:
            if (std::holds_alternative<std::string>(l)) return mismatch("string");
            if (std::holds_alternative<int64_t>(l)) f.rhs = (double)std::get<int64_t>(l);
            break;
        }
        return Plan{{f}};
    }
};
inline Expr col(const std::string& name) { return Expr::column(name); }
template <typename T> inline Expr lit(T v) { return Expr::literal(Expr::to_lit(v)); }

// ---- VecStore / VecQueryPlan (src/vec.rs) ---------------------------------------------------------------------------
class VecStore;
class VecQueryPlan {
    std::optional<std::vector<std::vector<float>>> queries_;
    std::optional<Metric> metric_;
    std::optional<std::pair<float, Cmp>> filter_;
    std::optional<TakeType> take_type_;
    std::optional<size_t> take_count_;
    const VecStore* store_ = nullptr;
    std::optional<std::vector<uint8_t>> row_mask_;

  public:
    VecQueryPlan() = default;
    static VecQueryPlan make() { return VecQueryPlan(); }
    VecQueryPlan with_vector_store(const VecStore& s) && { store_ = &s; return std::move(*this); }
    VecQueryPlan with_query_vectors(std::vector<std::vector<float>> q) && { queries_ = std::move(q); return std::move(*this); }
    VecQueryPlan with_metric(Metric m) && { metric_ = m; return std::move(*this); }
    VecQueryPlan with_row_mask(std::vector<uint8_t> m) && { row_mask_ = std::move(m); return std::move(*this); }
    VecQueryPlan filter(float score, Cmp c) && { filter_ = std::make_pair(score, c); return std::move(*this); }
    VecQueryPlan take(size_t n) && { return std::move(*this).take_with(n, std::nullopt); }
    VecQueryPlan take_min(size_t n) && { return std::move(*this).take_with(n, TakeType::Min); }
    VecQueryPlan take_max(size_t n) && { return std::move(*this).take_with(n, TakeType::Max); }
    Result<std::vector<SearchResult>> collect() &&;

  private:
    VecQueryPlan take_with(size_t n, std::optional<TakeType> tt) && {  // src/vec.rs:103-116
        take_count_ = n;
        if (tt) take_type_ = tt;
        else if (!take_type_ && metric_) take_type_ = infer_default_take_type(*metric_);
        return std::move(*this);
    }
};

class VecStore {
    size_t dim_;
    int32_t vector_format_ = OTTERS_VECTORS_FMT_F32;
    mutable otters_vecstore* h_ = nullptr;
    mutable std::vector<float> pending_;
    size_t n_ = 0;
    friend class VecQueryPlan;
    otters_vecstore* handle() const {
        if (!h_ && otters_vecstore_create_fmt(default_ctx(), (uint32_t)dim_, vector_format_, &h_) != OTTERS_OK) throw std::runtime_error(otters_last_error());
        if (!pending_.empty()) {
            if (otters_vecstore_add(h_, pending_.data(), pending_.size() / dim_) != OTTERS_OK) throw std::runtime_error(otters_last_error());
            pending_.clear();
        }
        return h_;
    }

  public:
    explicit VecStore(size_t dim, int32_t vector_format = OTTERS_VECTORS_FMT_F32) : dim_(dim), vector_format_(vector_format) {}
    VecStore(const VecStore&) = delete;
    ~VecStore() { if (h_) otters_vecstore_destroy(h_); }
    Result<bool> add_vector(const std::vector<float>& v) {  // src/vec.rs:357-371
        if (v.size() != dim_)
            return Result<bool>::Err("Input vector length " + std::to_string(v.size()) + " does not match expected dimension " + std::to_string(dim_));
        pending_.insert(pending_.end(), v.begin(), v.end());
        ++n_;
        return Result<bool>::Ok(true);
    }
    Result<bool> add_vectors(const std::vector<std::vector<float>>& vs) {
        for (auto& v : vs) { auto r = add_vector(v); if (r.is_err()) return r; }
        return Result<bool>::Ok(true);
    }
    size_t len() const { return n_; }
    bool is_empty() const { return n_ == 0; }
    size_t dim() const { return dim_; }
    VecQueryPlan query(std::vector<std::vector<float>> queries, Metric m) const {  // src/vec.rs:387-411
        return VecQueryPlan().with_vector_store(*this).with_query_vectors(std::move(queries)).with_metric(m);
    }
    VecQueryPlan query(std::vector<float> q, Metric m) const { return query(std::vector<std::vector<float>>{std::move(q)}, m); }
};

inline Result<std::vector<SearchResult>> VecQueryPlan::collect() && {
    using R = Result<std::vector<SearchResult>>;
    // validate (src/vec.rs:170-203)
    if (!queries_) return R::Err("Query vectors or their norms are not set");
    if (!metric_) return R::Err("Search metric is not set");
    if (!store_) return R::Err("Vector store is not set");
    if (queries_->empty()) return R::Err("No queries provided");
    for (auto& q : *queries_)
        if (q.size() != store_->dim_)
            return R::Err("Query vector length " + std::to_string(q.size()) + " does not match expected dimension " + std::to_string(store_->dim_));
    const size_t n = store_->len();
    const size_t k = take_count_.value_or(n);                 // src/vec.rs:213
    const TakeType tt = take_type_.value_or(TakeType::Max);   // src/vec.rs:214
    if (n == 0 || k == 0) return R::Ok({});
    std::vector<float> flat;
    for (auto& q : *queries_) flat.insert(flat.end(), q.begin(), q.end());
    std::vector<uint64_t> words;
    otters_vec_query vq{};
    vq.queries = flat.data(); vq.nq = (uint32_t)queries_->size(); vq.dim = (uint32_t)store_->dim_;
    vq.metric = (int32_t)*metric_; vq.take_type = (int32_t)tt; vq.k = k;
    if (filter_) { vq.has_filter = 1; vq.thr = filter_->first; vq.cmp = (int32_t)filter_->second; }
    if (row_mask_) {
        words.assign((row_mask_->size() + 63) / 64 + 1, 0);
        for (size_t i = 0; i < row_mask_->size(); ++i) if ((*row_mask_)[i]) words[i >> 6] |= 1ull << (i & 63);
        vq.row_mask_words = words.data(); vq.row_mask_bits = row_mask_->size();
    }
    const size_t cap = std::min(k, n * queries_->size());
    std::vector<uint64_t> idx(cap);
    std::vector<float> score(cap);
    uint64_t len = 0;
    if (otters_vecstore_query(store_->handle(), &vq, idx.data(), score.data(), nullptr, cap, &len) != OTTERS_OK) return R::Err(otters_last_error());
    std::vector<SearchResult> out;
    for (uint64_t i = 0; i < std::min<uint64_t>(len, cap); ++i) out.push_back({(size_t)idx[i], score[i]});
    return R::Ok(std::move(out));
}

// ---- MetaStore (src/meta.rs) -------------------------------------------------------------------------------------------
struct MetaQueryStats {  // src/meta.rs:832-842 (seconds)
    size_t total_chunks = 0, pruned_chunks = 0, evaluated_chunks = 0, vectors_compared = 0;
    double prune_duration = 0, score_duration = 0, merge_duration = 0, total_duration = 0;
};
struct MetaBuildStats { size_t n_rows = 0, dim = 0, n_chunks = 0; double vectors_ingest_duration = 0, zonemap_build_duration = 0, build_total_duration = 0; };
struct MetaQueryResults {  // src/meta.rs:23-40
    std::vector<std::string> columns;
    std::map<std::string, Column> data;
    std::vector<size_t> indices;
    std::vector<float> scores;
    size_t len() const { return indices.size(); }
    bool is_empty() const { return indices.empty(); }
};

class MetaStore;
class MetaQueryPlan {
    const MetaStore* store_;
    std::vector<std::vector<float>> queries_;
    Metric metric_;
    std::optional<CompiledFilter> filter_;
    std::optional<std::string> meta_error_;
    std::optional<std::pair<float, Cmp>> vec_filter_;
    std::optional<TakeType> take_type_;
    std::optional<size_t> take_count_;

  public:
    MetaQueryPlan(const MetaStore* s, std::vector<std::vector<float>> q, Metric m) : store_(s), queries_(std::move(q)), metric_(m) {}
    MetaQueryPlan meta_filter(const Expr& e) &&;
    MetaQueryPlan vec_filter(float score, Cmp c) && { vec_filter_ = std::make_pair(score, c); return std::move(*this); }
    MetaQueryPlan take(size_t k) && { take_count_ = k; take_type_ = infer_default_take_type(metric_); return std::move(*this); }  // src/meta.rs:623-630
    Result<MetaQueryResults> collect() &&;
};

class MetaStoreBuilder {
    std::map<std::string, DataType> schema_;
    std::vector<Column> cols_;
    std::optional<std::vector<std::vector<float>>> vectors_;
    size_t chunk_size_ = 1024;
    int bloom_mode_ = 0;
    double bloom_fpr_ = 0.01;
    uint64_t bloom_bits_ = 0;
    int32_t vector_format_ = OTTERS_VECTORS_FMT_F32;

  public:
    explicit MetaStoreBuilder(std::vector<Column> cols) : cols_(std::move(cols)) { for (auto& c : cols_) schema_[c.name()] = c.dtype(); }
    MetaStoreBuilder with_vectors(std::vector<std::vector<float>> v) && { vectors_ = std::move(v); return std::move(*this); }
    MetaStoreBuilder with_chunk_size(size_t c) && { chunk_size_ = std::max<size_t>(c, 1); return std::move(*this); }
    MetaStoreBuilder with_bloom_fpr(double f) && { bloom_mode_ = 0; bloom_fpr_ = std::isfinite(f) ? std::min(std::max(f, 1e-2), 0.5) : 0.01; return std::move(*this); }
    MetaStoreBuilder with_bloom_bits(size_t b) && { bloom_mode_ = 1; bloom_bits_ = std::max<size_t>(b, 64); return std::move(*this); }
    /* extension: OTTERS_VECTORS_FMT_BF16 keeps the rows as bf16 (see otters_vecstore_create_fmt in otters_b200.h) */
    MetaStoreBuilder with_vector_format(int32_t f) && { vector_format_ = f; return std::move(*this); }
    Result<std::unique_ptr<MetaStore>> build() &&;
};

class MetaStore {
    otters_metastore* h_ = nullptr;
    std::map<std::string, DataType> schema_;
    std::vector<Column> cols_;
    size_t chunk_size_ = 1024, dim_ = 0, n_rows_ = 0;
    mutable std::optional<MetaQueryStats> last_;
    std::optional<MetaBuildStats> build_stats_;
    friend class MetaStoreBuilder;
    friend class MetaQueryPlan;

  public:
    MetaStore() = default;
    MetaStore(const MetaStore&) = delete;
    ~MetaStore() { if (h_) otters_metastore_destroy(h_); }
    static MetaStoreBuilder from_columns(std::vector<Column> cols) { return MetaStoreBuilder(std::move(cols)); }
    const std::map<std::string, DataType>& schema() const { return schema_; }
    size_t n_chunks() const { return (size_t)otters_metastore_n_chunks(h_); }
    size_t chunk_size() const { return chunk_size_; }
    std::optional<MetaQueryStats> last_query_stats() const { return last_; }
    std::optional<MetaBuildStats> build_stats() const { return build_stats_; }
    MetaQueryPlan query(std::vector<float> q, Metric m) const { return MetaQueryPlan(this, {std::move(q)}, m); }
    MetaQueryPlan query_batch(std::vector<std::vector<float>> q, Metric m) const { return MetaQueryPlan(this, std::move(q), m); }
};

inline Result<std::unique_ptr<MetaStore>> MetaStoreBuilder::build() && {  // src/meta.rs:151-305
    using R = Result<std::unique_ptr<MetaStore>>;
    if (!vectors_) return R::Err("vectors must be provided to build MetaStore");
    const size_t n = vectors_->size();
    for (auto& c : cols_)
        if (c.len() != n)
            return R::Err("column '" + c.name() + "' length " + std::to_string(c.len()) + " does not match vectors length " + std::to_string(n));
    const size_t dim = n ? (*vectors_)[0].size() : 0;
    if (dim == 0 && n > 0) return R::Err("vector dimension cannot be zero");
    std::vector<float> flat;
    flat.reserve(n * dim);
    for (size_t i = 0; i < n; ++i) {
        if ((*vectors_)[i].size() != dim)
            return R::Err("vector at index " + std::to_string(i) + " has dim " + std::to_string((*vectors_)[i].size()) + ", expected " + std::to_string(dim));
        flat.insert(flat.end(), (*vectors_)[i].begin(), (*vectors_)[i].end());
    }
    std::vector<otters_column> cc(cols_.size());
    std::vector<std::vector<uint64_t>> nw(cols_.size()), offs(cols_.size());
    std::vector<std::vector<uint8_t>> bytes(cols_.size());
    for (size_t i = 0; i < cols_.size(); ++i) cols_[i].fill(&cc[i], &nw[i], &offs[i], &bytes[i]);
    otters_build_params bp{};
    bp.n_rows = n; bp.dim = (uint32_t)dim; bp.chunk_size = chunk_size_;
    bp.bloom_mode = bloom_mode_; bp.bloom_fpr = bloom_fpr_; bp.bloom_bits = bloom_bits_;
    bp.vectors_kind = OTTERS_VECTORS_HOST; bp.vectors = flat.data();
    bp.columns = cc.data(); bp.n_columns = (uint32_t)cc.size();
    bp.vector_format = vector_format_;
    auto ms = std::make_unique<MetaStore>();
    otters_build_stats bs{};
    if (otters_metastore_build(default_ctx(), &bp, &ms->h_, &bs) != OTTERS_OK) return R::Err(otters_last_error());
    ms->schema_ = schema_;
    ms->cols_ = std::move(cols_);
    ms->chunk_size_ = chunk_size_; ms->dim_ = dim; ms->n_rows_ = n;
    ms->build_stats_ = MetaBuildStats{(size_t)bs.n_rows, (size_t)bs.dim, (size_t)bs.n_chunks, bs.vectors_ingest_s, bs.zonemap_build_s, bs.build_total_s};
    return R::Ok(std::move(ms));
}

inline MetaQueryPlan MetaQueryPlan::meta_filter(const Expr& e) && {  // src/meta.rs:605-616
    auto r = e.compile(store_->schema_);
    if (r.is_ok()) { filter_ = r.unwrap(); meta_error_.reset(); }
    else meta_error_ = "meta_filter compile error: " + r.unwrap_err();
    return std::move(*this);
}

inline Result<MetaQueryResults> MetaQueryPlan::collect() && {  // src/meta.rs:632-829
    using R = Result<MetaQueryResults>;
    if (meta_error_) return R::Err(*meta_error_);
    const size_t n = store_->n_rows_;
    const size_t k = take_count_.value_or(n);
    const TakeType tt = take_type_.value_or(infer_default_take_type(metric_));
    bool bad = queries_.empty();
    for (auto& q : queries_) if (q.size() != queries_[0].size()) bad = true;
    std::vector<float> flat;
    if (!bad) for (auto& q : queries_) flat.insert(flat.end(), q.begin(), q.end());
    otters_vec_query vq{};
    vq.queries = bad ? nullptr : flat.data(); vq.nq = (uint32_t)queries_.size(); vq.dim = bad ? 0 : (uint32_t)queries_[0].size();
    vq.metric = (int32_t)metric_; vq.take_type = (int32_t)tt; vq.k = k;
    if (vec_filter_) { vq.has_filter = 1; vq.thr = vec_filter_->first; vq.cmp = (int32_t)vec_filter_->second; }
    std::vector<uint32_t> offs{0};
    std::vector<otters_leaf> leaves;
    std::vector<std::string> strs;
    otters_filter f{};
    if (filter_) {
        size_t nstr = 0;
        for (auto& cl : filter_->clauses) for (auto& lf : cl) if (std::holds_alternative<std::string>(lf.rhs)) ++nstr;
        strs.reserve(nstr);
        for (auto& cl : filter_->clauses) {
            for (auto& lf : cl) {
                otters_leaf L{};
                size_t ci = 0;
                while (ci < store_->cols_.size() && store_->cols_[ci].name() != lf.column) ++ci;
                L.col = (uint32_t)ci; L.op = (int32_t)lf.cmp;
                if (auto p = std::get_if<int64_t>(&lf.rhs)) { L.kind = OTTERS_LIT_I64; L.i = *p; }
                else if (auto p2 = std::get_if<double>(&lf.rhs)) { L.kind = OTTERS_LIT_F64; L.f = *p2; }
                else { strs.push_back(std::get<std::string>(lf.rhs)); L.kind = OTTERS_LIT_STR; L.s = (const uint8_t*)strs.back().data(); L.slen = strs.back().size(); }
                leaves.push_back(L);
            }
            offs.push_back((uint32_t)leaves.size());
        }
        f.n_clauses = (uint32_t)filter_->clauses.size(); f.clause_offsets = offs.data(); f.leaves = leaves.data();
    }
    const size_t cap = std::max<size_t>(std::min(k, n * std::max<size_t>(queries_.size(), 1)), 1);
    std::vector<uint64_t> idx(cap);
    std::vector<float> score(cap);
    uint64_t len = 0;
    otters_query_stats st{};
    if (otters_metastore_query(store_->h_, &vq, filter_ ? &f : nullptr, idx.data(), score.data(), nullptr, cap, &len, &st) != OTTERS_OK)
        return R::Err(otters_last_error());
    store_->last_ = MetaQueryStats{(size_t)st.total_chunks, (size_t)st.pruned_chunks, (size_t)st.evaluated_chunks, (size_t)st.vectors_compared,
                                   st.prune_s, st.score_s, st.merge_s, st.total_s};
    MetaQueryResults out;
    for (uint64_t i = 0; i < std::min<uint64_t>(len, cap); ++i) { out.indices.push_back((size_t)idx[i]); out.scores.push_back(score[i]); }
    for (auto& kv : store_->schema_) out.columns.push_back(kv.first);  // std::map iterates sorted (src/meta.rs:723-724)
    for (auto& c : store_->cols_) out.data.emplace(c.name(), c.gather(out.indices));
    return R::Ok(std::move(out));
}

}  // namespace otters
