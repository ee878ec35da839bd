#include "sliding_window_stream_to_pdu_impl.h"

#include <gnuradio/io_signature.h>

#include <stdexcept>

namespace gr {
namespace uwspr {

sliding_window_stream_to_pdu::sptr sliding_window_stream_to_pdu::make(int fs, int fl, int shift, int C)
{
    return gnuradio::get_initial_sptr(new sliding_window_stream_to_pdu_impl(fs, fl, shift, C));
}

sliding_window_stream_to_pdu_impl::sliding_window_stream_to_pdu_impl(int fs, int fl, int shift, int C)
    : gr::sync_block("sliding_window_stream_to_pdu", gr::io_signature::make(1, 1, sizeof(gr_complex)), gr::io_signature::make(0, 0, 0)),
      d_fl(fl), d_step(shift * fs), d_capacity((size_t)C * (size_t)fl), d_head(0), d_size(0), d_ready(0)
{
    // the reference peeks fl - shift*fs samples after popping shift*fs, so it needs 0 < shift*fs <= fl and a ring of
    // at least one window; it does not check
    if (fs <= 0 || fl <= 0 || d_step <= 0 || d_step > fl || C < 1)
        throw std::invalid_argument("uwspr.sliding_window_stream_to_pdu: need 0 < shift*fs <= fl and C >= 1");
    d_out_port = pmt::mp("out");
    message_port_register_out(d_out_port);
    d_ring.assign(d_capacity, gr_complex(0, 0));
}

sliding_window_stream_to_pdu_impl::~sliding_window_stream_to_pdu_impl() {}

int sliding_window_stream_to_pdu_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &)
{
    const gr_complex *in = static_cast<const gr_complex *>(input_items[0]);
    // a full ring drops its oldest sample for every new one (boost::circular_buffer::push_back)
    for (int i = 0; i < noutput_items; i++) {
        if (d_size == d_capacity) {
            d_ring[d_head] = in[i];
            d_head = (d_head + 1) % d_capacity;
        } else {
            d_ring[(d_head + d_size) % d_capacity] = in[i];
            d_size++;
        }
    }
    d_ready += noutput_items;
    if (d_ready >= d_fl) {
        // publish the oldest fl samples and retire the first shift*fs of them
        pmt::pmt_t vec = pmt::make_vector(d_fl, pmt::PMT_NIL);
        for (int i = 0; i < d_fl; i++) {
            const gr_complex &v = d_ring[(d_head + (size_t)i) % d_capacity];
            pmt::vector_set(vec, i, pmt::make_rectangular(v.real(), v.imag()));
        }
        d_head = (d_head + (size_t)d_step) % d_capacity;
        d_size -= (size_t)d_step;
        d_ready -= d_step;
        message_port_pub(d_out_port, pmt::cons(pmt::PMT_NIL, vec));
    }
    return noutput_items;
}

}  // namespace uwspr
}  // namespace gr
