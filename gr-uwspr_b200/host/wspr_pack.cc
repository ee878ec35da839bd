// Host side, transmit direction (test and simulation support): text -> 7-byte message ->
// 162 channel symbols.  The reference only ships the receive direction; these are the inverses
// of what it does, so that synthetic windows can carry real callsigns:
//   pack_type1       inverse of lib/helpers.cc:321-434 (unpack50 / unpackcall / unpackgrid) for
//                    type-1 messages "CALL GRID dBm"
//   channel_symbols  the convolutional encoder lib/Fano.cc:81-100 (polynomials :54-55), the
//                    inverse of the bit-reversal de-interleaver
//                    lib/sync_and_demodulate_impl.cc:265-282, and symbol = 2*data + pr3[i]
//                    (the relation lib/sync_and_demodulate_impl.cc:216-224 undoes)
#include <ctype.h>
#include <string.h>

#include "uwspr_b200.h"
#include "wspr_tables.h"

namespace {

// character codes of the callsign field: digits 0-9, letters 10-35, blank 36
int call_code(char ch)
{
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'A' && ch <= 'Z') return ch - 'A' + 10;
    if (ch == ' ') return 36;
    return -1;
}

inline unsigned parity32(unsigned v) { return (unsigned)__builtin_parity(v); }

}  // namespace

extern "C" int uwspr_b200_pack_type1(const char *call, const char *grid4, int dbm, int8_t *message7)
{
    if (!call || !grid4 || !message7) return UWSPR_B200_E_PARAM;
    // callsign: up to six characters with the digit in the third position (a call whose digit
    // is second, like "K1ABC", is shifted right by one blank)
    char c6[6] = { ' ', ' ', ' ', ' ', ' ', ' ' };
    const size_t len = strlen(call);
    if (len < 3 || len > 6) return UWSPR_B200_E_PARAM;
    char up[7];
    for (size_t i = 0; i < len; i++) up[i] = (char)toupper((unsigned char)call[i]);
    size_t at = 0;
    if (isdigit((unsigned char)up[2]))
        at = 0;
    else if (isdigit((unsigned char)up[1]) && len <= 5)
        at = 1;
    else
        return UWSPR_B200_E_PARAM;
    memcpy(c6 + at, up, len);
    int code[6];
    for (int i = 0; i < 6; i++) {
        code[i] = call_code(c6[i]);
        if (code[i] < 0) return UWSPR_B200_E_PARAM;
    }
    if (code[1] == 36 || code[2] > 9) return UWSPR_B200_E_PARAM;
    for (int i = 3; i < 6; i++)
        if (code[i] < 10) return UWSPR_B200_E_PARAM;  // letters or blanks only
    uint32_t n1 = (uint32_t)code[0];
    n1 = n1 * 36 + (uint32_t)code[1];
    n1 = n1 * 10 + (uint32_t)code[2];
    n1 = n1 * 27 + (uint32_t)(code[3] - 10);
    n1 = n1 * 27 + (uint32_t)(code[4] - 10);
    n1 = n1 * 27 + (uint32_t)(code[5] - 10);

    // locator: two letters A-R, two digits
    if (strlen(grid4) != 4) return UWSPR_B200_E_PARAM;
    const int g0 = toupper((unsigned char)grid4[0]) - 'A', g1 = toupper((unsigned char)grid4[1]) - 'A';
    const int g2 = grid4[2] - '0', g3 = grid4[3] - '0';
    if (g0 < 0 || g0 > 17 || g1 < 0 || g1 > 17 || g2 < 0 || g2 > 9 || g3 < 0 || g3 > 9) return UWSPR_B200_E_PARAM;
    const uint32_t ngrid = (uint32_t)((179 - 10 * g0 - g2) * 180 + 10 * g1 + g3);
    // power: type-1 messages carry 0..60 dBm ending in 0, 3 or 7
    if (dbm < 0 || dbm > 60 || !(dbm % 10 == 0 || dbm % 10 == 3 || dbm % 10 == 7)) return UWSPR_B200_E_PARAM;
    const uint32_t n2 = ngrid * 128 + (uint32_t)dbm + 64;

    message7[0] = (int8_t)(uint8_t)(n1 >> 20);
    message7[1] = (int8_t)(uint8_t)(n1 >> 12);
    message7[2] = (int8_t)(uint8_t)(n1 >> 4);
    message7[3] = (int8_t)(uint8_t)(((n1 & 0xfu) << 4) | ((n2 >> 18) & 0xfu));
    message7[4] = (int8_t)(uint8_t)(n2 >> 10);
    message7[5] = (int8_t)(uint8_t)(n2 >> 2);
    message7[6] = (int8_t)(uint8_t)((n2 & 3u) << 6);
    return UWSPR_B200_OK;
}

extern "C" void uwspr_b200_channel_symbols(const int8_t *message7, uint8_t *symbols162)
{
    // 50 payload bits + 31 zero tail bits through the K=32, r=1/2 encoder: two output bits per
    // input bit, most significant bit of every byte first
    uint8_t bits[176];
    uint8_t data[11];
    memset(data, 0, sizeof(data));
    memcpy(data, message7, 7);
    unsigned state = 0;
    int n = 0;
    for (int b = 0; b < 11; b++)
        for (int k = 7; k >= 0; k--) {
            state = (state << 1) | ((data[b] >> k) & 1u);
            bits[n++] = (uint8_t)parity32(state & 0xf2d05351u);
            bits[n++] = (uint8_t)parity32(state & 0xe4613c47u);
        }
    // interleave: encoder output p goes to position bitrev8(i) for the p-th i in 0..255 whose
    // bit-reversed value is below 162
    uint8_t inter[UWSPR_B200_NSYM];
    int p = 0;
    for (int i = 0; i < 256 && p < UWSPR_B200_NSYM; i++) {
        unsigned r = (unsigned)i;
        r = ((r & 0xf0u) >> 4) | ((r & 0x0fu) << 4);
        r = ((r & 0xccu) >> 2) | ((r & 0x33u) << 2);
        r = ((r & 0xaau) >> 1) | ((r & 0x55u) << 1);
        if (r < (unsigned)UWSPR_B200_NSYM) inter[r] = bits[p++];
    }
    for (int i = 0; i < UWSPR_B200_NSYM; i++) {
        const unsigned sync = (WSPR_SYNC_WORDS[i >> 5] >> (i & 31)) & 1u;
        symbols162[i] = (uint8_t)(2u * inter[i] + sync);
    }
}
