// Host side, upstream of the hot path: the .c2 sample file reader of uwspr.c2file_source.
//
// Reference behaviour being matched: lib/c2file_source_impl.cc:75-96 -- a 14-byte name, an int
// (WSPR type / minutes), a double (dial frequency, MHz), then 45000 interleaved (I, Q) fp32
// samples at 375 sps; the block hands out I - jQ (the quadrature sign is flipped at :91) and
// refuses files with fewer samples.
#include <stdio.h>
#include <string.h>

#include <vector>

#include "uwspr_b200.h"

extern "C" int uwspr_b200_read_c2(const char *path, float *iq, char *name15, int32_t *type, double *freq_mhz)
{
    if (!path || !iq) return UWSPR_B200_E_PARAM;
    FILE *fp = fopen(path, "rb");
    if (!fp) return UWSPR_B200_E_PARAM;   // the reference throws "can't open file"
    char name[14];
    int32_t ntrmin = 0;
    double dfreq = 0.0;
    const size_t npts = 45000;
    std::vector<float> buffer(2 * npts);
    bool ok = fread(name, 1, 14, fp) == 14 && fread(&ntrmin, sizeof(ntrmin), 1, fp) == 1 &&
              fread(&dfreq, sizeof(dfreq), 1, fp) == 1;
    ok = ok && fread(buffer.data(), sizeof(float), 2 * npts, fp) == 2 * npts;  // "invalid number of samples"
    fclose(fp);
    if (!ok) return UWSPR_B200_E_PARAM;
    for (size_t i = 0; i < npts; i++) {
        iq[2 * i] = buffer[2 * i];
        iq[2 * i + 1] = -buffer[2 * i + 1];
    }
    if (name15) {
        memcpy(name15, name, 14);
        name15[14] = 0;
    }
    if (type) *type = ntrmin;
    if (freq_mhz) *freq_mhz = dfreq;
    return UWSPR_B200_OK;
}
