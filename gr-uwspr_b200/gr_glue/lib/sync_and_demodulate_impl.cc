#include "sync_and_demodulate_impl.h"

#include <gnuradio/io_signature.h>

#include <stdexcept>
#include <string>

#include "pdu_codec.h"

namespace gr {
namespace uwspr {

sync_and_demodulate::sptr sync_and_demodulate::make(int fs, int fl, int spb, int maxdrift, int maxfreqs, int cf)
{
    return gnuradio::get_initial_sptr(new sync_and_demodulate_impl(fs, fl, spb, maxdrift, maxfreqs, cf));
}

sync_and_demodulate_impl::sync_and_demodulate_impl(int fs, int fl, int spb, int maxdrift, int maxfreqs, int cf)
    : gr::block("sync_and_demodulate", gr::io_signature::make(0, 0, 0), gr::io_signature::make(0, 0, 0)),
      d_ctx(NULL), d_fl(fl), d_maxfreqs(maxfreqs), d_framecount(0), d_iq(NULL), d_log(NULL)
{
    d_in_port = pmt::mp("in");
    message_port_register_in(d_in_port);
    set_msg_handler(d_in_port, boost::bind(&sync_and_demodulate_impl::demodulate, this, _1));
    d_out_port = pmt::mp("out");
    message_port_register_out(d_out_port);

    // half pass bandwidth and threshold belong to FDR; the fine stage does not read them
    uwspr_b200_params_t p = { fs, fl, spb, maxdrift, maxfreqs, 10, cf, 10, 0, 1, maxfreqs, 0 };
    const int st = uwspr_b200_create(&p, &d_ctx);
    if (st != UWSPR_B200_OK) throw std::invalid_argument(std::string("uwspr.sync_and_demodulate: ") + uwspr_b200_create_error());
    void *a = NULL;
    if (uwspr_b200_host_alloc(&a, sizeof(float) * 2 * (size_t)fl) != UWSPR_B200_OK) {
        uwspr_b200_destroy(d_ctx);
        throw std::runtime_error("uwspr.sync_and_demodulate: cannot allocate pinned host buffers");
    }
    d_iq = static_cast<float *>(a);
    d_cands.resize(maxfreqs);
    d_retry.resize(maxfreqs);
    d_refined.resize(maxfreqs);
    d_jig.resize((size_t)maxfreqs * UWSPR_B200_NJIG);
    d_soft.resize((size_t)maxfreqs * UWSPR_B200_NJIG * UWSPR_B200_NSYM);
    // message log in the working directory, opened for append, with a start stamp (reference :98-108)
    d_log = fopen("messagelog.txt", "a");
    if (d_log)
        fprintf(stderr, "Messages logged in file messagelog.txt\n");
    else
        fprintf(stderr, "Error opening message log file!\n");
    time(&d_start);
    if (d_log) {
        fprintf(d_log, "Start time: %s\n", asctime(localtime(&d_start)));
        fflush(d_log);
    }
}

sync_and_demodulate_impl::~sync_and_demodulate_impl()
{
    if (d_log) fclose(d_log);
    uwspr_b200_host_free(d_iq);
    uwspr_b200_destroy(d_ctx);
}

// two clock lines, then the frame text (reference printtime() :300-313 and :508-525)
void sync_and_demodulate_impl::log_frame(const uwspr_b200_candidate_t &cand, const int8_t message7[7])
{
    if (!d_log) return;
    time_t now;
    time(&now);
    fprintf(d_log, "Handoff time : %s", asctime(localtime(&now)));
    const long dt = (long)difftime(now, d_start);
    fprintf(d_log, "Elapsed time: %02d:%02d:%02d\n", (int)((dt / 3600) % 24), (int)((dt / 60) % 60), (int)(dt % 60));
    char text[512];
    if (uwspr_b200_format_message_log(d_framecount, &cand, message7, text, sizeof(text)) == UWSPR_B200_OK) fputs(text, d_log);
    fflush(d_log);
}

void sync_and_demodulate_impl::demodulate(pmt::pmt_t msg)
{
    const pmt::pmt_t tuple(pmt::cdr(msg));
    glue::window_from_pmt(pmt::tuple_ref(tuple, 0), d_fl, d_iq);
    const int npk = (int)pmt::to_long(pmt::tuple_ref(tuple, 1));
    if (npk < 0 || npk > d_maxfreqs) throw std::runtime_error("uwspr.sync_and_demodulate: candidate count outside [0, maxfreqs]");
    const pmt::pmt_t list(pmt::tuple_ref(tuple, 2));
    for (int i = 0; i < npk; i++) d_cands[i] = glue::candidate_from_pmt(pmt::vector_ref(list, i));
    if (npk == 0) return;

    // Stage 1: refinement chain + the soft symbols of the first jiggle of every candidate (idt = 0, :460-475).
    // Clean frames decode there, so the other sixteen jiggles are evaluated only for the candidates that did not.
    int32_t n32 = npk;
    int st = uwspr_b200_fine(d_ctx, d_iq, UWSPR_B200_HOST, d_fl, 1, &n32, d_cands.data(), npk, 0, 1, d_refined.data(),
                             d_jig.data(), d_soft.data());
    if (st != UWSPR_B200_OK) throw std::runtime_error(std::string("uwspr.sync_and_demodulate: ") + uwspr_b200_last_error(d_ctx));
    std::vector<int8_t> messages((size_t)npk * 7, 0);
    std::vector<char> decoded(npk, 0);
    std::vector<int> retry;
    for (int j = 0; j < npk; j++) {
        if (uwspr_b200_decode_candidate(&d_refined[j], &d_jig[j], &d_soft[(size_t)j * UWSPR_B200_NSYM], 1, &messages[7 * (size_t)j], NULL, NULL))
            decoded[j] = 1;
        else if (d_refined[j].worth_a_try)
            retry.push_back(j);
    }
    if (!retry.empty()) {
        // Stage 2: jiggles idt = 1..16 of the candidates still undecoded, in one submission
        const int nr = (int)retry.size(), nj = UWSPR_B200_NJIG - 1;
        for (int r = 0; r < nr; r++) d_retry[r] = d_cands[retry[r]];
        n32 = nr;
        st = uwspr_b200_fine(d_ctx, d_iq, UWSPR_B200_HOST, d_fl, 1, &n32, d_retry.data(), nr, 1, nj, d_refined.data(), d_jig.data(),
                             d_soft.data());
        if (st != UWSPR_B200_OK) throw std::runtime_error(std::string("uwspr.sync_and_demodulate: ") + uwspr_b200_last_error(d_ctx));
        for (int r = 0; r < nr; r++) {
            const int j = retry[r];
            if (uwspr_b200_decode_candidate(&d_refined[r], &d_jig[(size_t)r * nj], &d_soft[(size_t)r * nj * UWSPR_B200_NSYM], nj,
                                            &messages[7 * (size_t)j], NULL, NULL))
                decoded[j] = 1;
        }
    }
    // one PDU per decoded candidate, in candidate order, no de-duplication (reference :484-531)
    for (int j = 0; j < npk; j++) {
        if (!decoded[j]) continue;
        d_framecount++;
        log_frame(d_cands[j], &messages[7 * (size_t)j]);
        message_port_pub(d_out_port, glue::message_pdu(&messages[7 * (size_t)j]));
    }
}

}  // namespace uwspr
}  // namespace gr
