/* Oracle stub for <pmt/pmt.h> (test infrastructure, not product code).
 *
 * A small polymorphic-value type with the subset of the PMT API the reference
 * blocks call.  Values are immutable-by-convention and reference counted, as in
 * GNU Radio.  Type errors throw std::runtime_error (GNU Radio throws
 * pmt::wrong_type). */
#ifndef ORACLE_STUB_PMT_H
#define ORACLE_STUB_PMT_H

#include <complex>
#include <cstddef>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pmt {

struct pmt_base {
    enum kind_t { K_NIL, K_SYMBOL, K_LONG, K_DOUBLE, K_COMPLEX, K_PAIR, K_VECTOR, K_TUPLE, K_BLOB };
    kind_t kind;
    explicit pmt_base(kind_t k) : kind(k) {}
    virtual ~pmt_base() {}
};
typedef std::shared_ptr<pmt_base> pmt_t;

struct pmt_symbol : pmt_base { std::string s; pmt_symbol(const std::string &v) : pmt_base(K_SYMBOL), s(v) {} };
struct pmt_long : pmt_base { long v; pmt_long(long x) : pmt_base(K_LONG), v(x) {} };
struct pmt_double : pmt_base { double v; pmt_double(double x) : pmt_base(K_DOUBLE), v(x) {} };
struct pmt_complex : pmt_base { std::complex<double> v; pmt_complex(std::complex<double> x) : pmt_base(K_COMPLEX), v(x) {} };
struct pmt_pair : pmt_base { pmt_t a, d; pmt_pair(pmt_t x, pmt_t y) : pmt_base(K_PAIR), a(x), d(y) {} };
struct pmt_vector : pmt_base { std::vector<pmt_t> v; pmt_vector(size_t n, pmt_t fill) : pmt_base(K_VECTOR), v(n, fill) {} };
struct pmt_tuple : pmt_base { std::vector<pmt_t> v; pmt_tuple() : pmt_base(K_TUPLE) {} };
struct pmt_blob : pmt_base { std::vector<unsigned char> v; pmt_blob(const void *p, size_t n) : pmt_base(K_BLOB), v((const unsigned char *)p, (const unsigned char *)p + n) {} };

static const pmt_t PMT_NIL(new pmt_base(pmt_base::K_NIL));

template <class T> inline T *pmt_cast(const pmt_t &p, pmt_base::kind_t k, const char *what)
{
    if (!p || p->kind != k) throw std::runtime_error(std::string("pmt wrong_type: ") + what);
    return static_cast<T *>(p.get());
}

inline pmt_t mp(const std::string &s) { return pmt_t(new pmt_symbol(s)); }
inline pmt_t mp(const char *s) { return pmt_t(new pmt_symbol(s)); }
inline pmt_t intern(const std::string &s) { return mp(s); }
inline std::string symbol_to_string(const pmt_t &p) { return pmt_cast<pmt_symbol>(p, pmt_base::K_SYMBOL, "symbol")->s; }

inline pmt_t from_long(long v) { return pmt_t(new pmt_long(v)); }
inline long to_long(const pmt_t &p) { return pmt_cast<pmt_long>(p, pmt_base::K_LONG, "long")->v; }
inline pmt_t from_double(double v) { return pmt_t(new pmt_double(v)); }
inline double to_double(const pmt_t &p)
{
    if (p && p->kind == pmt_base::K_LONG) return (double)static_cast<pmt_long *>(p.get())->v;
    return pmt_cast<pmt_double>(p, pmt_base::K_DOUBLE, "double")->v;
}
inline pmt_t from_complex(std::complex<double> v) { return pmt_t(new pmt_complex(v)); }
inline pmt_t make_rectangular(double re, double im) { return pmt_t(new pmt_complex(std::complex<double>(re, im))); }
inline std::complex<double> to_complex(const pmt_t &p) { return pmt_cast<pmt_complex>(p, pmt_base::K_COMPLEX, "complex")->v; }

inline pmt_t cons(const pmt_t &a, const pmt_t &d) { return pmt_t(new pmt_pair(a, d)); }
inline pmt_t car(const pmt_t &p) { return pmt_cast<pmt_pair>(p, pmt_base::K_PAIR, "pair")->a; }
inline pmt_t cdr(const pmt_t &p) { return pmt_cast<pmt_pair>(p, pmt_base::K_PAIR, "pair")->d; }

inline pmt_t make_vector(size_t n, const pmt_t &fill) { return pmt_t(new pmt_vector(n, fill)); }
inline pmt_t vector_ref(const pmt_t &p, size_t k)
{
    pmt_vector *v = pmt_cast<pmt_vector>(p, pmt_base::K_VECTOR, "vector");
    if (k >= v->v.size()) throw std::out_of_range("pmt vector_ref");
    return v->v[k];
}
inline void vector_set(const pmt_t &p, size_t k, const pmt_t &x)
{
    pmt_vector *v = pmt_cast<pmt_vector>(p, pmt_base::K_VECTOR, "vector");
    if (k >= v->v.size()) throw std::out_of_range("pmt vector_set");
    v->v[k] = x;
}
inline size_t length(const pmt_t &p)
{
    if (p && p->kind == pmt_base::K_VECTOR) return static_cast<pmt_vector *>(p.get())->v.size();
    if (p && p->kind == pmt_base::K_TUPLE) return static_cast<pmt_tuple *>(p.get())->v.size();
    if (p && p->kind == pmt_base::K_BLOB) return static_cast<pmt_blob *>(p.get())->v.size();
    throw std::runtime_error("pmt wrong_type: length");
}

template <class... A> inline pmt_t make_tuple(const A &... a)
{
    pmt_tuple *t = new pmt_tuple();
    pmt_t r(t);
    const pmt_t items[] = { a... };
    t->v.assign(items, items + sizeof...(a));
    return r;
}
inline pmt_t tuple_ref(const pmt_t &p, size_t k)
{
    pmt_tuple *t = pmt_cast<pmt_tuple>(p, pmt_base::K_TUPLE, "tuple");
    if (k >= t->v.size()) throw std::out_of_range("pmt tuple_ref");
    return t->v[k];
}

inline pmt_t make_blob(const void *buf, size_t len) { return pmt_t(new pmt_blob(buf, len)); }
inline size_t blob_length(const pmt_t &p) { return pmt_cast<pmt_blob>(p, pmt_base::K_BLOB, "blob")->v.size(); }
inline const void *blob_data(const pmt_t &p) { return pmt_cast<pmt_blob>(p, pmt_base::K_BLOB, "blob")->v.data(); }
inline const void *uniform_vector_elements(const pmt_t &p, size_t &len)
{
    pmt_blob *b = pmt_cast<pmt_blob>(p, pmt_base::K_BLOB, "blob");
    len = b->v.size();
    return b->v.data();
}
inline bool is_null(const pmt_t &p) { return p && p->kind == pmt_base::K_NIL; }

} // namespace pmt
#endif
