// pmt.h -- minimal stand-in for GNU Radio's polymorphic types, just the subset gr-amps uses
// (blob, tuple, long, uint64, bool, symbol, pair, u8vector, dict).  Only compiled when the real
// <pmt/pmt.h> is not available; the block sources are written against the real API names.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pmt {

struct pmt_base {
    enum kind_t { NIL, SYMBOL, LONG, UINT64, BOOL, BLOB, TUPLE, PAIR, U8VECTOR, DICT } kind = NIL;
    std::string sym;
    long l = 0;
    uint64_t u = 0;
    bool b = false;
    std::vector<uint8_t> bytes;
    std::vector<std::shared_ptr<pmt_base>> items;
};
typedef std::shared_ptr<pmt_base> pmt_t;

inline pmt_t make(pmt_base::kind_t k) { pmt_t p(new pmt_base()); p->kind = k; return p; }
static const pmt_t PMT_NIL = make(pmt_base::NIL);

inline pmt_t intern(const std::string &s) { pmt_t p = make(pmt_base::SYMBOL); p->sym = s; return p; }
inline pmt_t mp(const std::string &s) { return intern(s); }
inline pmt_t mp(const char *s) { return intern(s); }
inline std::string symbol_to_string(const pmt_t &p) { return p->sym; }
inline pmt_t make_blob(const void *d, size_t n) {
    pmt_t p = make(pmt_base::BLOB);
    p->bytes.assign(static_cast<const uint8_t *>(d), static_cast<const uint8_t *>(d) + n);
    return p;
}
inline pmt_t mp(const void *d, size_t n) { return make_blob(d, n); }
inline bool is_blob(const pmt_t &p) { return p->kind == pmt_base::BLOB; }
inline const void *blob_data(const pmt_t &p) { return p->bytes.data(); }
inline size_t blob_length(const pmt_t &p) { return p->bytes.size(); }
inline pmt_t from_long(long v) { pmt_t p = make(pmt_base::LONG); p->l = v; return p; }
inline long to_long(const pmt_t &p) { if (p->kind != pmt_base::LONG) throw std::runtime_error("pmt: not a long"); return p->l; }
inline pmt_t from_uint64(uint64_t v) { pmt_t p = make(pmt_base::UINT64); p->u = v; return p; }
inline uint64_t to_uint64(const pmt_t &p) { return p->kind == pmt_base::UINT64 ? p->u : (uint64_t)p->l; }
inline pmt_t from_bool(bool v) { pmt_t p = make(pmt_base::BOOL); p->b = v; return p; }
inline bool to_bool(const pmt_t &p) { return p->b; }
inline bool is_tuple(const pmt_t &p) { return p->kind == pmt_base::TUPLE; }
inline pmt_t make_tuple_v(const std::vector<pmt_t> &v) { pmt_t p = make(pmt_base::TUPLE); p->items = v; return p; }
template <typename... A> pmt_t make_tuple(const A &...a) { return make_tuple_v(std::vector<pmt_t>{a...}); }
inline pmt_t tuple_ref(const pmt_t &p, size_t i) { return p->items.at(i); }
inline size_t length(const pmt_t &p) { return p->kind == pmt_base::TUPLE ? p->items.size() : p->bytes.size(); }
inline pmt_t make_dict() { return make(pmt_base::DICT); }
inline pmt_t cons(const pmt_t &a, const pmt_t &d) { pmt_t p = make(pmt_base::PAIR); p->items = {a, d}; return p; }
inline pmt_t car(const pmt_t &p) { return p->items.at(0); }
inline pmt_t cdr(const pmt_t &p) { return p->items.at(1); }
inline bool is_pair(const pmt_t &p) { return p->kind == pmt_base::PAIR; }
inline pmt_t init_u8vector(size_t n, const uint8_t *d) { pmt_t p = make(pmt_base::U8VECTOR); p->bytes.assign(d, d + n); return p; }
inline const uint8_t *u8vector_elements(const pmt_t &p, size_t &n) { n = p->bytes.size(); return p->bytes.data(); }
inline std::vector<uint8_t> u8vector_elements(const pmt_t &p) { return p->bytes; }
inline bool is_u8vector(const pmt_t &p) { return p->kind == pmt_base::U8VECTOR; }

}  // namespace pmt
