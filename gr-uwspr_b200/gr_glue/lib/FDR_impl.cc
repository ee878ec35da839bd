#include "FDR_impl.h"

#include <gnuradio/io_signature.h>

#include <stdexcept>
#include <string>

#include "pdu_codec.h"

namespace gr {
namespace uwspr {

FDR::sptr FDR::make(int fs, int fl, int spb, int maxdrift, int maxfreqs, int halfbandwidth, int cf, int threshold)
{
    return gnuradio::get_initial_sptr(new FDR_impl(fs, fl, spb, maxdrift, maxfreqs, halfbandwidth, cf, threshold));
}

FDR_impl::FDR_impl(int fs, int fl, int spb, int maxdrift, int maxfreqs, int halfbandwidth, int cf, int threshold)
    : gr::block("FDR", gr::io_signature::make(0, 0, 0), gr::io_signature::make(0, 0, 0)),
      d_ctx(NULL), d_fl(fl), d_maxfreqs(maxfreqs), d_iq(NULL), d_cands(NULL)
{
    d_in_port = pmt::mp("in");
    message_port_register_in(d_in_port);
    set_msg_handler(d_in_port, boost::bind(&FDR_impl::transform, this, _1));
    d_out_port = pmt::mp("out");
    message_port_register_out(d_out_port);

    uwspr_b200_params_t p = { fs, fl, spb, maxdrift, maxfreqs, halfbandwidth, cf, threshold, /*device*/ 0,
                              /*max_windows*/ 1, /*max_candidates*/ maxfreqs, /*nonlinear_intended_t*/ 0 };
    // The reference prints and exit()s on a half pass bandwidth above fs/2 (lib/FDR_impl.cc:82-90) and reads out of
    // bounds a little below that; here every parameter outside the supported domain is a constructor exception.
    const int st = uwspr_b200_create(&p, &d_ctx);
    if (st != UWSPR_B200_OK) throw std::invalid_argument(std::string("uwspr.FDR: ") + uwspr_b200_create_error());
    void *a = NULL, *b = NULL;
    if (uwspr_b200_host_alloc(&a, sizeof(float) * 2 * (size_t)fl) != UWSPR_B200_OK ||
        uwspr_b200_host_alloc(&b, sizeof(uwspr_b200_candidate_t) * (size_t)maxfreqs) != UWSPR_B200_OK) {
        uwspr_b200_host_free(a);
        uwspr_b200_destroy(d_ctx);
        throw std::runtime_error("uwspr.FDR: cannot allocate pinned host buffers");
    }
    d_iq = static_cast<float *>(a);
    d_cands = static_cast<uwspr_b200_candidate_t *>(b);
}

FDR_impl::~FDR_impl()
{
    uwspr_b200_host_free(d_iq);
    uwspr_b200_host_free(d_cands);
    uwspr_b200_destroy(d_ctx);
}

void FDR_impl::transform(pmt::pmt_t msg)
{
    const pmt::pmt_t window(pmt::cdr(msg));   // car is the (empty) metadata
    glue::window_from_pmt(window, d_fl, d_iq);
    int32_t npk = 0, total = 0;
    const int st = uwspr_b200_coarse(d_ctx, d_iq, UWSPR_B200_HOST, d_fl, 1, &npk, d_cands, d_maxfreqs, &total);
    if (st != UWSPR_B200_OK) throw std::runtime_error(std::string("uwspr.FDR: ") + uwspr_b200_last_error(d_ctx));
    message_port_pub(d_out_port, glue::candidates_pdu(window, d_cands, npk));
}

}  // namespace uwspr
}  // namespace gr
