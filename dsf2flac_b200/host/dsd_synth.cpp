// dsd_synth.cpp -- synthetic 1-bit DSD test signals: a 2nd-order
// sigma-delta modulator driven by a sine, packed 8 samples per byte.  Used by tests/ and
// bench.py to make inputs with real DSD byte statistics; not part of the decimation path.
#include <cmath>
#include <cstdint>

extern "C" {

// v = sign(s2); s1 += x - v; s2 += s1 - v, state 0, x[n] = amp*sin(2*pi*freq*n/fs).
// lsb_first != 0: the first sample in time goes to bit 0 of its byte (the .dsf packing, which
// the reference flags msbIsPlayedFirst()==true); else to bit 7 (the .dff packing).
void dsdhost_synth_sdm(uint8_t* out, int64_t nbytes, double fs, double freq, double amp, int lsb_first) {
    double s1 = 0.0, s2 = 0.0;
    // rotate a unit phasor instead of calling sin() per sample; renormalise now and then
    const double w = 2.0 * M_PI * freq / fs;
    const double cw = std::cos(w), sw = std::sin(w);
    double re = 1.0, im = 0.0;
    for (int64_t k = 0; k < nbytes; ++k) {
        unsigned byte = 0;
        for (int b = 0; b < 8; ++b) {
            const double x = amp * im;
            const double v = s2 >= 0.0 ? 1.0 : -1.0;
            s1 += x - v;
            s2 += s1 - v;
            if (v > 0) byte |= lsb_first ? (1u << b) : (0x80u >> b);
            const double nre = re * cw - im * sw;
            im = re * sw + im * cw;
            re = nre;
        }
        if ((k & 0xfff) == 0xfff) {
            const double n = 1.0 / std::sqrt(re * re + im * im);
            re *= n; im *= n;
        }
        out[k] = (uint8_t)byte;
    }
}

}  // extern "C"
