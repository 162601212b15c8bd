// Smoke test of the C++ facade (include/otters.hpp): a few of the reference's own tests, written the way the
// reference writes them (tests/vec_store_tests.rs, tests/meta_zonemap_tests.rs, tests/expr_tests.rs).
#include <cmath>
#include <cstdio>
#include <set>

#include "otters.hpp"

using namespace otters;

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

static void host_only_tests() {
    // tests/expr_tests.rs:201-208 — tautology dropped
    std::map<std::string, DataType> schema{{"age", DataType::Int64}, {"score", DataType::Float64}, {"name", DataType::String}, {"ts", DataType::DateTime}};
    auto cf = ((col("name").eq("bob") | col("name").neq("bob")) & col("age").gt(5)).compile(schema);
    CHECK(cf.is_ok() && cf.unwrap().clauses.size() == 1 && cf.unwrap().clauses[0][0].column == "age");
    // :149-161 — A & (B | C) => [[A],[B,C]]
    auto cnf = (col("age").gt(25) & (col("score").gte(80.0) | col("age").lt(18))).compile(schema);
    CHECK(cnf.is_ok() && cnf.unwrap().clauses.size() == 2 && cnf.unwrap().clauses[1].size() == 2);
    // :91-102 — type mismatch
    auto bad = col("age").gt(25.5).compile(schema);
    CHECK(bad.is_err() && bad.unwrap_err() == "Type mismatch for column 'age': expected Int64, got literal float");
    // :170-190 — datetime literal
    auto dt = col("ts").gte("2023-01-02T03:04:05Z").compile(schema);
    CHECK(dt.is_ok() && std::get<int64_t>(dt.unwrap().clauses[0][0].rhs) == 1672628645000LL);
    CHECK(parse_datetime_millis("2024-01-01").value() == 1704067200000LL);
    CHECK(parse_datetime_millis("2024-12-31 23:59:59").value() == 1735689599000LL);
    CHECK(!parse_datetime_millis("nonsense").has_value());
    // tests/vec_store_tests.rs:51-63 — deferred dimension error, no device needed
    VecStore store(3);
    CHECK(store.add_vector({1.0f, 0.0f, 0.0f}).is_ok());
    CHECK(store.add_vector({1.0f, 2.0f}).is_err());
    auto r = store.query(std::vector<float>{1.0f, 0.0f}, Metric::Cosine).take(5).collect();
    CHECK(r.is_err() && r.unwrap_err().find("Query vector length 2 does not match expected dimension 3") != std::string::npos);
    CHECK(VecQueryPlan().take(5).collect().unwrap_err() == "Query vectors or their norms are not set");
}

static void device_tests() {
    {  // tests/vec_store_tests.rs:251-276 — dot product ranking
        VecStore store(2);
        store.add_vectors({{3.0f, 4.0f}, {1.0f, 1.0f}, {0.0f, 1.0f}, {-1.0f, 0.0f}});
        auto res = store.query(std::vector<float>{3.0f, 4.0f}, Metric::DotProduct).take(4).collect().unwrap();
        CHECK(res.size() == 4 && res[0].score == 25.0f && res[3].score == -3.0f && res[0].index == 0);
    }
    {  // :323-343 — take_min
        VecStore store(2);
        store.add_vectors({{1.0f, 0.0f}, {2.0f, 0.0f}, {0.5f, 0.0f}, {-1.0f, 0.0f}});
        auto res = store.query(std::vector<float>{1.0f, 0.0f}, Metric::DotProduct).take_min(2).collect().unwrap();
        CHECK(res.size() == 2 && res[0].score == -1.0f && res[1].score == 0.5f);
        auto flt = store.query(std::vector<float>{1.0f, 0.0f}, Metric::DotProduct).filter(1.0f, Cmp::Gt).take(10).collect().unwrap();
        CHECK(flt.size() == 1 && flt[0].score == 2.0f);
    }
    {  // tests/meta_zonemap_tests.rs:17-89,133-156
        std::vector<std::vector<float>> vectors(9, {1.0f, 0.0f});
        using OI = std::optional<int32_t>;
        using OS = std::optional<std::string>;
        auto val = Column("val", DataType::Int32).from(std::vector<OI>{1, 2, std::nullopt, 10, 11, 12, std::nullopt, std::nullopt, std::nullopt});
        auto ts = Column("ts", DataType::DateTime).from(std::vector<OS>{OS("2024-01-01T00:00:00Z"), std::nullopt, OS("2024-06-01T00:00:00Z"), OS("2026-01-01T00:00:00Z"),
                                                                       OS("2026-06-01T00:00:00Z"), OS("2024-12-31T23:59:59Z"), std::nullopt, std::nullopt, std::nullopt});
        auto grade = Column("grade", DataType::String).from(std::vector<OS>{OS("A"), OS("B"), std::nullopt, OS("C"), OS("A"), OS("A"), std::nullopt, std::nullopt, std::nullopt});
        std::vector<Column> cols;
        cols.push_back(std::move(val));
        cols.push_back(std::move(ts));
        cols.push_back(std::move(grade));
        auto built = MetaStore::from_columns(std::move(cols)).with_vectors(vectors).with_chunk_size(3).build();
        CHECK(built.is_ok());
        auto& store = *built.unwrap();
        auto res = store.query({1.0f, 0.0f}, Metric::DotProduct).meta_filter(col("val").gt(5)).take(9).collect().unwrap();
        std::set<size_t> got(res.indices.begin(), res.indices.end());
        CHECK((got == std::set<size_t>{3, 4, 5}));
        auto st = store.last_query_stats().value();
        CHECK(st.total_chunks == 3 && st.evaluated_chunks == 1 && st.pruned_chunks == 2);
        auto res2 = store.query({1.0f, 0.0f}, Metric::DotProduct).meta_filter(col("val").gt(5) & col("ts").lt("2025-01-01T00:00:00Z")).take(9).collect().unwrap();
        CHECK(res2.len() == 1 && res2.indices[0] == 5);
        CHECK(res2.columns == (std::vector<std::string>{"grade", "ts", "val"}));
        CHECK(res2.data.at("val").i32_values()[0] == 12);
        auto err = store.query({1.0f, 0.0f}, Metric::Cosine).meta_filter(col("val").gt(1.5)).take(3).collect();
        CHECK(err.is_err() && err.unwrap_err().rfind("meta_filter compile error: Type mismatch for column 'val'", 0) == 0);
    }
}

int main(int argc, char** argv) {
    host_only_tests();
    bool device = argc > 1 && std::string(argv[1]) == "--device";
    if (device) device_tests();
    if (failures) std::printf("FACADE_TEST_FAILED (%d)\n", failures);
    else std::printf("FACADE_TEST_OK%s\n", device ? " (host+device)" : " (host only)");
    return failures ? 1 : 0;
}
