/* Oracle stub for <gnuradio/block.h> (test infrastructure, not product code).
 *
 * Just enough of gr::block for the reference's message-driven blocks to be
 * constructed and driven synchronously from a harness:
 *   - message ports are names; set_msg_handler() stores the functor,
 *   - message_port_pub() appends to a per-block outbox the harness drains,
 *   - oracle_deliver() invokes the stored handler on the calling thread.
 * There is no scheduler, no threads and no flowgraph. */
#ifndef ORACLE_STUB_GR_BLOCK_H
#define ORACLE_STUB_GR_BLOCK_H

#include <complex>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <pmt/pmt.h>

typedef std::complex<float> gr_complex;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;
typedef std::vector<int> gr_vector_int;

namespace boost {
using std::shared_ptr;
template <class F, class... A>
auto bind(F &&f, A &&... a) -> decltype(std::bind(std::forward<F>(f), std::forward<A>(a)...))
{
    return std::bind(std::forward<F>(f), std::forward<A>(a)...);
}
} // namespace boost
using std::placeholders::_1;
using std::placeholders::_2;

namespace gr {

class io_signature
{
public:
    typedef boost::shared_ptr<io_signature> sptr;
    int min_streams, max_streams, item_size;
    static sptr make(int mn, int mx, int sz)
    {
        sptr s(new io_signature());
        s->min_streams = mn;
        s->max_streams = mx;
        s->item_size = sz;
        return s;
    }
};

class block
{
public:
    block() {}
    block(const std::string &name, io_signature::sptr, io_signature::sptr) : d_name(name) {}
    virtual ~block() {}

    void message_port_register_in(pmt::pmt_t) {}
    void message_port_register_out(pmt::pmt_t) {}
    template <class H> void set_msg_handler(pmt::pmt_t port, H h)
    {
        d_handlers[pmt::symbol_to_string(port)] = std::function<void(pmt::pmt_t)>(h);
    }
    void message_port_pub(pmt::pmt_t port, pmt::pmt_t msg)
    {
        (void)port;
        d_outbox.push_back(msg);
    }
    void set_alignment(int) {}
    const std::string &name() const { return d_name; }

    /* harness side */
    void oracle_deliver(const std::string &port, pmt::pmt_t msg) { d_handlers.at(port)(msg); }
    std::deque<pmt::pmt_t> &oracle_outbox() { return d_outbox; }

private:
    std::string d_name;
    std::map<std::string, std::function<void(pmt::pmt_t)>> d_handlers;
    std::deque<pmt::pmt_t> d_outbox;
};

} // namespace gr

namespace gnuradio {
template <class T> boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); }
} // namespace gnuradio

#endif
