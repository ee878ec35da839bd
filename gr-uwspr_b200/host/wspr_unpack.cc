// Host side, downstream of the hot path: the 7-byte decoded blob -> "CALL GRID dBm" text that
// uwspr.WSPR_unpacker publishes.
//
// Reference behaviour being matched (file:line under the reference tree):
//   lib/helpers.cc:321-353   unpack50: 28-bit n1 (callsign), 22-bit n2 (locator + power/type)
//   lib/helpers.cc:355-401   unpackcall
//   lib/helpers.cc:403-434   unpackgrid
//   lib/helpers.cc:436-492   unpackpfx (type 2: prefix / suffix)
//   lib/helpers.cc:151-319   nhash = Bob Jenkins' lookup3 hashlittle(), masked to 15 bits
//   lib/helpers.cc:494-590   unpk_: type 1 / 2 / 3 dispatch and the callsign hash table
#include <ctype.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "uwspr_b200.h"

namespace {

const char kAlphabet[] = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ ";

inline uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

// lookup3 hashlittle() of a short key (byte-wise tail handling), & 32767
uint32_t callsign_hash(const char *key, size_t length, uint32_t initval)
{
    uint32_t a, b, c;
    a = b = c = 0xdeadbeefu + (uint32_t)length + initval;
    const unsigned char *k = reinterpret_cast<const unsigned char *>(key);
    while (length > 12) {
        a += k[0] | (uint32_t)k[1] << 8 | (uint32_t)k[2] << 16 | (uint32_t)k[3] << 24;
        b += k[4] | (uint32_t)k[5] << 8 | (uint32_t)k[6] << 16 | (uint32_t)k[7] << 24;
        c += k[8] | (uint32_t)k[9] << 8 | (uint32_t)k[10] << 16 | (uint32_t)k[11] << 24;
        a -= c; a ^= rotl(c, 4); c += b;
        b -= a; b ^= rotl(a, 6); a += c;
        c -= b; c ^= rotl(b, 8); b += a;
        a -= c; a ^= rotl(c, 16); c += b;
        b -= a; b ^= rotl(a, 19); a += c;
        c -= b; c ^= rotl(b, 4); b += a;
        length -= 12;
        k += 12;
    }
    if (length == 0) return c;  // lookup3 returns before the final mix (and before the mask)
    uint32_t w[3] = { 0, 0, 0 };
    for (size_t i = 0; i < length; i++) w[i >> 2] += (uint32_t)k[i] << (8 * (i & 3));
    a += w[0];
    b += w[1];
    c += w[2];
    c ^= b; c -= rotl(b, 14);
    a ^= c; a -= rotl(c, 11);
    b ^= a; b -= rotl(a, 25);
    c ^= b; c -= rotl(b, 16);
    a ^= c; a -= rotl(c, 4);
    b ^= a; b -= rotl(a, 14);
    c ^= b; c -= rotl(b, 24);
    return c & 32767u;
}

// 28-bit callsign field -> up to six characters; false when the field is out of range.
// The reference formats the characters left-justified in a six-wide field and then turns every
// blank into a terminator; `sixth` receives what is left at offset 5 of that buffer (0 if blank),
// which the type-3 path reads back as the first locator character.
bool decode_call(int32_t n, std::string &call, char *sixth = nullptr)
{
    if (n >= 262177560) return false;
    char t[7];
    t[5] = kAlphabet[n % 27 + 10]; n /= 27;
    t[4] = kAlphabet[n % 27 + 10]; n /= 27;
    t[3] = kAlphabet[n % 27 + 10]; n /= 27;
    t[2] = kAlphabet[n % 10]; n /= 10;
    t[1] = kAlphabet[n % 36]; n /= 36;
    t[0] = kAlphabet[n];
    t[6] = 0;
    int i = 0;
    while (i < 5 && t[i] == ' ') i++;   // leading blanks go (at most five are skipped)
    char padded[8];
    snprintf(padded, sizeof(padded), "%-6s", &t[i]);
    if (sixth) *sixth = (padded[5] == ' ') ? 0 : padded[5];
    call.assign(padded);
    const size_t cut = call.find(' ');
    if (cut != std::string::npos) call.resize(cut);
    return true;
}

bool decode_grid(int32_t n2, char grid[5])
{
    const int32_t ngrid = n2 >> 7;
    if (ngrid >= 32400) return false;
    int dlat = (ngrid % 180) - 90;
    int dlong = (ngrid / 180) * 2 - 180 + 2;
    if (dlong < -180) dlong += 360;
    if (dlong > 180) dlong += 360;
    const int nlong = (int)(60.0 * (180.0 - dlong) / 5.0);
    int n1 = nlong / 240, n2b = (nlong - 240 * n1) / 24;
    grid[0] = kAlphabet[10 + n1];
    grid[2] = kAlphabet[n2b];
    const int nlat = (int)(60.0 * (dlat + 90) / 2.5);
    n1 = nlat / 240;
    n2b = (nlat - 240 * n1) / 24;
    grid[1] = kAlphabet[10 + n1];
    grid[3] = kAlphabet[n2b];
    grid[4] = 0;
    return true;
}

// type 2: add a 1-3 character prefix or a 1-2 character suffix to the callsign
bool apply_prefix(int32_t nprefix, std::string &call)
{
    if (nprefix < 60000) {
        char pfx[4] = { 0, 0, 0, 0 };
        int32_t n = nprefix;
        for (int i = 2; i >= 0; i--) {
            const int nc = n % 37;
            pfx[i] = (nc <= 9) ? (char)(nc + '0') : (nc <= 35) ? (char)(nc + 'A' - 10) : ' ';
            n /= 37;
        }
        const char *p = strrchr(pfx, ' ');
        call = std::string(p ? p + 1 : pfx) + "/" + call;
        return true;
    }
    // the reference narrows the suffix code to a (signed) char before testing it
    const int nc = (int)(signed char)(nprefix - 60000);
    if (nc >= 0 && nc <= 9) {
        call += std::string("/") + (char)(nc + '0');
    } else if (nc >= 10 && nc <= 35) {
        call += std::string("/") + (char)(nc + 'A' - 10);
    } else if (nc >= 36 && nc <= 125) {
        call += std::string("/") + (char)((nc - 26) / 10 + '0') + (char)((nc - 26) % 10 + '0');
    } else {
        return false;
    }
    return true;
}

}  // namespace

extern "C" {

size_t uwspr_b200_hashtab_bytes(void) { return 32768 * 13; }

int uwspr_b200_unpack(const int8_t *message7, char *hashtab, char *text, size_t text_cap)
{
    if (!message7 || !hashtab || !text || text_cap < 24) return -1;
    text[0] = 0;
    const unsigned char *dat = reinterpret_cast<const unsigned char *>(message7);
    // 50 payload bits, MSB first: n1 = bits 0..27, n2 = bits 28..49
    const int32_t n1 = (int32_t)dat[0] << 20 | (int32_t)dat[1] << 12 | (int32_t)dat[2] << 4 | (dat[3] >> 4);
    const int32_t n2 = (int32_t)(dat[3] & 15) << 18 | (int32_t)dat[4] << 10 | (int32_t)dat[5] << 2 | (dat[6] >> 6);
    std::string call;
    char grid[5];
    if (!decode_call(n1, call)) return 1;
    if (!decode_grid(n2, grid)) return 1;
    const int ntype = (n2 & 127) - 64;
    int noprint = 0;
    char cdbm[16];
    if (ntype >= 0 && ntype <= 62) {
        const int nu = ntype % 10;
        if (nu == 0 || nu == 3 || nu == 7) {
            // type 1: callsign, 4-character locator, power
            snprintf(cdbm, sizeof(cdbm), "%2d", ntype);
            snprintf(text, text_cap, "%s %.4s %.2s", call.c_str(), grid, cdbm);
            strcpy(hashtab + callsign_hash(call.c_str(), call.size(), 146) * 13, call.c_str());
        } else {
            // type 2: compound callsign, power
            int nadd = nu;
            if (nu > 3) nadd = nu - 3;
            if (nu > 7) nadd = nu - 7;
            const int32_t n3 = n2 / 128 + 32768 * (nadd - 1);
            if (!apply_prefix(n3, call)) return 1;
            const int ndbm = ntype - nadd;
            snprintf(cdbm, sizeof(cdbm), "%2d", ndbm);
            snprintf(text, text_cap, "%s %.2s", call.c_str(), cdbm);
            const int nu2 = ndbm % 10;
            if (nu2 == 0 || nu2 == 3 || nu2 == 7 || nu2 == 10) {
                if (call.size() <= 12) strcpy(hashtab + callsign_hash(call.c_str(), call.size(), 146) * 13, call.c_str());
            } else {
                noprint = 1;
            }
        }
    } else if (ntype < 0) {
        // type 3: hashed callsign, 6-character locator (carried in the callsign field, rotated), power
        const int ndbm = -(ntype + 1);
        std::string grid6;
        char last = 0;
        decode_call(n1, call, &last);
        if (last) grid6 += last;
        grid6 += call.substr(0, 5);
        const int nu = ndbm % 10;
        const unsigned char g0 = grid6.size() > 0 ? grid6[0] : 0, g1 = grid6.size() > 1 ? grid6[1] : 0;
        const unsigned char g2 = grid6.size() > 2 ? grid6[2] : 0, g3 = grid6.size() > 3 ? grid6[3] : 0;
        if ((nu != 0 && nu != 3 && nu != 7 && nu != 10) || !isalpha(g0) || !isalpha(g1) || !isdigit(g2) || !isdigit(g3))
            noprint = 1;
        const int ihash = (n2 - ntype - 64) / 128;
        std::string shown;
        if (hashtab[ihash * 13] != 0)
            shown = std::string("<") + (hashtab + ihash * 13) + ">";
        else
            shown = "<...>";
        snprintf(cdbm, sizeof(cdbm), "%2d", ndbm);
        snprintf(text, text_cap, "%s %s %.2s", shown.c_str(), grid6.c_str(), cdbm);
        if (ntype == -64) noprint = 1;
    }
    return noprint;
}

}  // extern "C"
