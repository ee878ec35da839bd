// Drives the GNU Radio glue of gr-uwspr_b200/gr_glue (gr::block subclasses with the reference's factories,
// ports and PDU schemas) against the minimal GNU Radio / PMT stand-ins of oracle/stubs.
//
//   test_gr_glue sliding
//       CPU only: the sliding-window block fed in uneven pieces; window k == stream[k*shift*fs, +fl),
//       at most one PDU per work() call.
//   test_gr_glue chain <window.npy> <out.bin>
//       one window PDU -> FDR -> sync_and_demodulate; writes npk, the candidate tuples as candidate_t
//       records and the published blobs to out.bin for the Python test to compare with the reference.
//   test_gr_glue stream <stream.npy> <shift_seconds> <out.bin>
//       the three blocks wired as in the example flowgraphs.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <complex>
#include <string>
#include <vector>

#include <uwspr/FDR.h>
#include <uwspr/sliding_window_stream_to_pdu.h>
#include <uwspr/sync_and_demodulate.h>

#include "../../gr-uwspr_b200/gr_glue/lib/pdu_codec.h"

using namespace gr::uwspr;

static std::vector<std::complex<float>> load_npy_c64(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    unsigned char h[10];
    if (fread(h, 1, 10, f) != 10 || memcmp(h, "\x93NUMPY", 6) != 0) { fprintf(stderr, "not an npy file\n"); exit(2); }
    size_t hlen = h[8] | (h[9] << 8);
    std::string hdr(hlen, ' ');
    if (fread(&hdr[0], 1, hlen, f) != hlen || hdr.find("<c8") == std::string::npos) { fprintf(stderr, "need complex64 npy\n"); exit(2); }
    std::vector<std::complex<float>> v;
    std::complex<float> buf[4096];
    size_t n;
    while ((n = fread(buf, sizeof(buf[0]), 4096, f)) > 0) v.insert(v.end(), buf, buf + n);
    fclose(f);
    return v;
}

static pmt::pmt_t window_pdu(const std::complex<float> *x, int fl)
{
    pmt::pmt_t vec = pmt::make_vector(fl, pmt::PMT_NIL);
    for (int i = 0; i < fl; i++) pmt::vector_set(vec, i, pmt::make_rectangular(x[i].real(), x[i].imag()));
    return pmt::cons(pmt::PMT_NIL, vec);
}

struct Dump {
    FILE *f;
    void i32(int32_t v) { fwrite(&v, 4, 1, f); }
};

// FDR -> sync_and_demodulate for one window PDU; appends [npk][npk x candidate_t][nblob][nblob x 7 bytes]
static int run_chain(FDR::sptr fdr, sync_and_demodulate::sptr sd, pmt::pmt_t pdu, Dump &d)
{
    fdr->oracle_deliver("in", pdu);
    if (fdr->oracle_outbox().size() != 1) { fprintf(stderr, "FDR published %zu PDUs\n", fdr->oracle_outbox().size()); return 1; }
    pmt::pmt_t out = fdr->oracle_outbox().front();
    fdr->oracle_outbox().pop_front();
    if (!pmt::is_null(pmt::car(out))) { fprintf(stderr, "car is not PMT_NIL\n"); return 1; }
    pmt::pmt_t tuple = pmt::cdr(out);
    if (pmt::tuple_ref(tuple, 0).get() != pmt::cdr(pdu).get()) { fprintf(stderr, "the window vector was copied, not forwarded\n"); return 1; }
    const int npk = (int)pmt::to_long(pmt::tuple_ref(tuple, 1));
    pmt::pmt_t list = pmt::tuple_ref(tuple, 2);
    if ((int)pmt::length(list) != npk) { fprintf(stderr, "candidate vector length != npk\n"); return 1; }
    d.i32(npk);
    for (int i = 0; i < npk; i++) {
        pmt::pmt_t t = pmt::vector_ref(list, i);
        const size_t want = pmt::to_long(pmt::tuple_ref(t, 0)) == 0 ? 6 : 9;
        if (pmt::length(t) != want) { fprintf(stderr, "candidate tuple arity %zu\n", pmt::length(t)); return 1; }
        uwspr_b200_candidate_t c = glue::candidate_from_pmt(t);
        fwrite(&c, sizeof(c), 1, d.f);
    }
    sd->oracle_deliver("in", out);
    d.i32((int32_t)sd->oracle_outbox().size());
    while (!sd->oracle_outbox().empty()) {
        pmt::pmt_t m = sd->oracle_outbox().front();
        sd->oracle_outbox().pop_front();
        if (!pmt::is_null(pmt::car(m)) || pmt::blob_length(pmt::cdr(m)) != 7) { fprintf(stderr, "bad message PDU\n"); return 1; }
        fwrite(pmt::blob_data(pmt::cdr(m)), 1, 7, d.f);
    }
    return 0;
}

static int test_sliding()
{
    const int fs = 375, fl = 45000, shift = 9, step = shift * fs;
    sliding_window_stream_to_pdu::sptr sw = sliding_window_stream_to_pdu::make(fs, fl, shift, 2);
    const int total = fl + 5 * step + 1234;
    std::vector<std::complex<float>> s(total);
    for (int i = 0; i < total; i++) s[i] = std::complex<float>((float)i, (float)(-2 * i));
    int pos = 0, k = 0, bad = 0, multi = 0;
    gr_vector_void_star none;
    const int pieces[] = { 1000, 4096, 777, 8191, 3000, 1, 2500 };
    for (int it = 0; pos < total; it++) {
        const int n = std::min(pieces[it % 7], total - pos);
        gr_vector_const_void_star in(1, &s[pos]);
        if (sw->work(n, in, none) != n) bad++;
        pos += n;
        if (sw->oracle_outbox().size() > 1) multi++;
        while (!sw->oracle_outbox().empty()) {
            pmt::pmt_t pdu = sw->oracle_outbox().front();
            sw->oracle_outbox().pop_front();
            pmt::pmt_t vec = pmt::cdr(pdu);
            if (!pmt::is_null(pmt::car(pdu)) || (int)pmt::length(vec) != fl) bad++;
            for (int i = 0; i < fl; i += 97) {
                const std::complex<double> c = pmt::to_complex(pmt::vector_ref(vec, i));
                if (c.real() != (double)(float)(k * step + i) || c.imag() != (double)(float)(-2 * (k * step + i))) bad++;
            }
            k++;
        }
    }
    printf("windows %d bad %d multi %d\n", k, bad, multi);
    bool threw = false;
    try { sliding_window_stream_to_pdu::make(fs, fl, 200, 2); } catch (const std::invalid_argument &) { threw = true; }
    printf("rejects shift*fs > fl: %d\n", (int)threw);
    return bad || multi || !threw;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !strcmp(argv[1], "sliding")) return test_sliding();
    const int fs = 375, fl = 45000, spb = 256, maxfreqs = 200, hbw = 10, cf = 1500, thr = 10;
    const int maxdrift = getenv("GLUE_MAXDRIFT") ? atoi(getenv("GLUE_MAXDRIFT")) : 0;
    if (argc >= 4 && !strcmp(argv[1], "chain")) {
        std::vector<std::complex<float>> x = load_npy_c64(argv[2]);
        if ((int)x.size() != fl) { fprintf(stderr, "window must hold %d samples\n", fl); return 2; }
        Dump d = { fopen(argv[3], "wb") };
        FDR::sptr fdr = FDR::make(fs, fl, spb, maxdrift, maxfreqs, hbw, cf, thr);
        sync_and_demodulate::sptr sd = sync_and_demodulate::make(fs, fl, spb, maxdrift, maxfreqs, cf);
        int rc = run_chain(fdr, sd, window_pdu(x.data(), fl), d);
        fclose(d.f);
        bool threw = false;   // the reference exits on this one (lib/FDR_impl.cc:82-90)
        try { FDR::make(fs, fl, spb, 0, maxfreqs, 188, cf, thr); } catch (const std::invalid_argument &) { threw = true; }
        printf("chain rc %d frames %d rejects halfbandwidth 188: %d\n", rc, 0, (int)threw);
        return rc || !threw;
    }
    if (argc >= 5 && !strcmp(argv[1], "stream")) {
        std::vector<std::complex<float>> s = load_npy_c64(argv[2]);
        const int shift = atoi(argv[3]);
        Dump d = { fopen(argv[4], "wb") };
        sliding_window_stream_to_pdu::sptr sw = sliding_window_stream_to_pdu::make(fs, fl, shift, 2);
        FDR::sptr fdr = FDR::make(fs, fl, spb, maxdrift, maxfreqs, hbw, cf, thr);
        sync_and_demodulate::sptr sd = sync_and_demodulate::make(fs, fl, spb, maxdrift, maxfreqs, cf);
        gr_vector_void_star none;
        int windows = 0, rc = 0;
        for (size_t pos = 0; pos < s.size() && !rc; pos += 4096) {
            const int n = (int)std::min<size_t>(4096, s.size() - pos);
            gr_vector_const_void_star in(1, &s[pos]);
            sw->work(n, in, none);
            while (!sw->oracle_outbox().empty() && !rc) {
                pmt::pmt_t pdu = sw->oracle_outbox().front();
                sw->oracle_outbox().pop_front();
                rc = run_chain(fdr, sd, pdu, d);
                windows++;
            }
        }
        fclose(d.f);
        printf("stream rc %d windows %d\n", rc, windows);
        return rc;
    }
    fprintf(stderr, "usage: test_gr_glue sliding | chain <window.npy> <out.bin> | stream <stream.npy> <shift> <out.bin>\n");
    return 2;
}
