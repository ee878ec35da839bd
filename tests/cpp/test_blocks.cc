// Drives the host-side mirrors of the three reference blocks (gr-uwspr_b200/host/blocks.h) the way a
// flowgraph wires them: sliding_window_stream_to_pdu -> FDR -> sync_and_demodulate.
//
//   test_blocks window <file.npy>   one 45000-sample window through FDR -> sync_and_demodulate
//                                   (needs a GPU); prints one line per decoded message: hex blob
//   test_blocks sliding             the sliding window alone (CPU only): checks that window k is
//                                   stream[k*shift*fs, k*shift*fs + fl) and at most one PDU per work()
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "blocks.h"

using namespace gr::uwspr;

static std::vector<gr_complex> load_npy(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    char magic[10];
    f.read(magic, 10);
    const unsigned hlen = (unsigned char)magic[8] | ((unsigned char)magic[9] << 8);
    f.seekg(10 + hlen);
    std::vector<gr_complex> x(45000);
    f.read(reinterpret_cast<char *>(x.data()), sizeof(gr_complex) * x.size());
    if (!f) throw std::runtime_error("short .npy file");
    return x;
}

static int run_window(const char *path)
{
    auto fdr = FDR::make(375, 45000, 256, 0, 200, 10, 1500, 10);
    auto sd = sync_and_demodulate::make(375, 45000, 256, 0, 200, 1500);
    fdr->set_msg_out([&](const candidates_pdu &p) {
        std::printf("candidates %zu\n", p.candidates.size());
        sd->demodulate(p);
    });
    sd->set_msg_out([&](const message_pdu &m) {
        std::printf("message ");
        for (int i = 0; i < 7; i++) std::printf("%02x", (unsigned)(unsigned char)m.blob[i]);
        std::printf("\n");
    });
    samples_ptr w(new std::vector<gr_complex>(load_npy(path)));
    fdr->transform(w);
    std::printf("frames %d\n", sd->framecount());
    return 0;
}

static int run_sliding()
{
    const int fs = 375, fl = 45000, shift = 9, C = 2;
    auto sw = sliding_window_stream_to_pdu::make(fs, fl, shift, C);
    std::vector<gr_complex> stream(fl + 5 * shift * fs);
    for (size_t i = 0; i < stream.size(); i++) stream[i] = gr_complex((float)i, -(float)(i % 977));
    int emitted = 0, bad = 0, calls_with_two = 0;
    sw->set_msg_out([&](samples_ptr w) {
        const size_t start = (size_t)emitted * shift * fs;
        if ((int)w->size() != fl) bad++;
        for (int i = 0; i < fl; i++)
            if ((*w)[i] != stream[start + i]) { bad++; break; }
        emitted++;
    });
    size_t pos = 0;
    while (pos < stream.size()) {
        const int n = (int)std::min<size_t>(1125, stream.size() - pos);
        const int before = emitted;
        if (sw->work(n, stream.data() + pos) != n) bad++;
        if (emitted - before > 1) calls_with_two++;
        pos += n;
    }
    std::printf("windows %d bad %d multi %d\n", emitted, bad, calls_with_two);
    return (emitted == 6 && bad == 0 && calls_with_two == 0) ? 0 : 1;
}

int main(int argc, char **argv)
{
    try {
        if (argc >= 3 && !std::strcmp(argv[1], "window")) return run_window(argv[2]);
        if (argc >= 2 && !std::strcmp(argv[1], "sliding")) return run_sliding();
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    std::fprintf(stderr, "usage: test_blocks window <file.npy> | sliding\n");
    return 64;
}
