/*
 * lte_b200.h -- C ABI of liblte_b200.so, the B200 (sm_100a) native engine for the
 * LTE simulate-and-count-BER link chain.
 *
 * The reference (Darioxavierl/OFDM-LTE) is pure Python/NumPy and has no FFI of its
 * own; each entry point below replaces the NumPy body of the reference function
 * cited next to it (paths relative to the reference root).  INTEGRATION.md shows
 * the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *  - Every function returns LTE_OK (0) or a negative LTE_ERR_* code; nothing throws.
 *  - All data pointers are DEVICE pointers owned by the caller unless the name ends
 *    in _host.  The library allocates nothing but the per-plan constant tables at
 *    lte_plan_create: kernels that need scratch take a caller-owned `workspace` whose
 *    size comes from the matching *_workspace_bytes() query (16-byte aligned, contents
 *    undefined afterwards).  No entry point allocates, frees or synchronises, so one plan
 *    may be used from several CUDA streams at once (with one workspace per stream).
 *  - No environment variable changes what a kernel computes.
 *  - Random draws are Philox2x32-10 keyed (seed, stream or row id, sample); ids are 32-bit counter words.  A call
 *    whose ids (id0 + count) would pass 2^32 returns LTE_ERR_UNSUPPORTED instead of silently repeating draws:
 *    use a new seed for the next 2^32 streams.
 *  - Launches are asynchronous on `stream` (a cudaStream_t passed as void*).
 *  - Complex samples are interleaved float pairs (re, im): `lte_c32`.
 *  - A "stream" is one link realisation: S OFDM symbols back to back, L = N + cp
 *    samples each.  Batches are stream-major: [B][...].
 *  - QAM symbols travel as one byte per symbol holding the natural-binary,
 *    MSB-first constellation index (core/modulator.py:80-86).
 *  - Frequency-domain tensors cover the bin window [k0, k0 + nk) of the N-point
 *    grid (raw FFT-bin order, no fftshift -- core/resource_mapper.py:45-74);
 *    LTE_WINDOW_FULL selects k0 = 0, nk = N, LTE_WINDOW_USEFUL the Nc occupied bins.
 */
#ifndef LTE_B200_H
#define LTE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTE_OK 0
#define LTE_ERR_INVALID_ARG (-1)
#define LTE_ERR_UNSUPPORTED (-2)
#define LTE_ERR_CUDA (-3)
#define LTE_ERR_NO_DEVICE (-4)

#define LTE_MAX_TAPS 8      /* config.py:34-60: longest ITU-R M.1225 profile has 8 taps */
#define LTE_JAKES_TONES 16  /* core/rayleighchannel.py:20 (N_s = 16) */
#define LTE_SLOT_SYMBOLS 14 /* core/lte_receiver.py:233 */
#define LTE_MAX_RX 8
#define LTE_MAX_TX 8
#define LTE_MAX_LAYERS 4

#define LTE_DET_MMSE 0
#define LTE_DET_ZF 1
#define LTE_DET_SIC 2
#define LTE_DET_MRC 3

#define LTE_BF_MRT 0       /* BeamformingPrecoder.calculate_mrt_weights (update_mode='adaptive') */
#define LTE_BF_CODEBOOK 1  /* LTECodebook.select_best_pmi precoder (update_mode='static') */

#define LTE_BLK_COLS 8     /* int32 columns of the code-block layout table (lte_tb_encode / lte_tb_decode) */

#define LTE_WINDOW_FULL 0
#define LTE_WINDOW_USEFUL 1

typedef struct { float re, im; } lte_c32;
typedef struct lte_plan lte_plan;

/* Numerology of one link (config.py:101-130) plus the pilot table.
 * pilots_host: [num_tx_pilot_sets][num_pilots] complex values, zero where a TX
 * antenna does not own the pilot (core/resource_mapper.py:137-152,
 * core/mimo_channel_estimator_periodic.py:75-106).  The host generates them
 * because they come from NumPy's MT19937 stream. */
typedef struct {
    int32_t N;                 /* FFT size, power of two, 64..2048 */
    int32_t Nc;                /* occupied subcarriers */
    int32_t cp;                /* cyclic-prefix samples */
    int32_t bits_per_symbol;   /* 2, 4 or 6 */
    int32_t mode_simple;       /* 1: mode='simple' (symbols on bins 0..Nc-1, no pilots) */
    int32_t num_tx_pilot_sets; /* 1 for SISO/SIMO */
    double  fs;                /* sampling rate in Hz (config.py:114) */
} lte_plan_desc;

/* Tapped-delay-line description after the host-side table work
 * (core/channel.py:162-186, core/rayleighchannel.py:16,52). */
typedef struct {
    int32_t num_taps;                 /* 0 => AWGN only (no fading) */
    int32_t delay[LTE_MAX_TAPS];      /* int(round(delay_s * fs)) */
    float   gain[LTE_MAX_TAPS];       /* faithful (double-converted) linear gains */
    double  doppler_hz;               /* fD = v * fc / c (core/channel.py:141-143) */
} lte_channel_desc;

int lte_version(void);               /* 100 * major + 10 * minor; 200 = caller-owned workspaces, spectral link, compact sweep layout */
const char* lte_error_string(int code);

/* Plan = device-resident constant tables: bin classes, pilot values, constellation
 * levels, slicer thresholds, FFT twiddles.  Created on the current CUDA device. */
int lte_plan_create(const lte_plan_desc* desc, const lte_c32* pilots_host, lte_plan** out);
int lte_plan_destroy(lte_plan* plan);
int lte_plan_num_data(const lte_plan* plan);    /* Nd (core/resource_mapper.py:80-83) */
int lte_plan_num_pilots(const lte_plan* plan);  /* Np (core/resource_mapper.py:85-88) */
/* copies the data / pilot bin indices to host arrays of Nd / Np int32 */
int lte_plan_indices_host(const lte_plan* plan, int32_t* data_idx_host, int32_t* pilot_idx_host);
int lte_plan_window(const lte_plan* plan, int window, int32_t* k0, int32_t* nk);

/* --- bit <-> symbol-index helpers (core/modulator.py:74-84, :109-110) ---------- */
/* bits: [B][nbits] bytes holding 0/1; idx: [B][nsym]; bits past nbits read as 0.
 * lte_bits_to_indices only: nbits < 0 means each row is np.packbits() output
 * (MSB-first bytes, ceil(-nbits/8) bytes per row) holding -nbits bits. */
int lte_bits_to_indices(const lte_plan*, const uint8_t* bits, int64_t nbits, uint8_t* idx,
                        int64_t nsym, int32_t B, void* stream);
int lte_indices_to_bits(const lte_plan*, const uint8_t* idx, int64_t nsym, uint8_t* bits,
                        int64_t nbits, int32_t B, void* stream);

/* constellation[idx] alone (QAMModulator.bits_to_symbols, core/modulator.py:84-86): the input of
 * the SC-FDM precoder.  idx, symbols: n elements. */
int lte_qam_map(const lte_plan*, const uint8_t* idx, lte_c32* symbols, int64_t n, void* stream);

/* --- stage 1+2 TX: QAM map + resource grid + IFFT*sqrt(N) + CP -------------------
 * replaces QAMModulator.bits_to_symbols (core/modulator.py:61-88),
 * ResourceMapper.map_symbols (core/resource_mapper.py:181-223) and
 * OFDMModulator._modulate_lte / modulate_stream (core/modulator.py:214-302).
 * idx: [B][S][Nd] symbol indices, or (symbols != NULL) pre-computed complex data
 * symbols [B*T][S][Nd] (SC-FDM precoded, SFBC/SM encoded), antenna t using pilot
 * set t.  tx: [B*T][S*L].  qam_out (optional): [B][S][Nd] mapped constellation
 * points.  stats (optional): [B*T][2] doubles = {max |x|^2, sum |x|^2}, accumulated with
 * atomics, caller zeroes (OFDMTransmitter.calculate_papr, core/ofdm_core.py:114-147). */
int lte_tx_map_ifft(const lte_plan*, const uint8_t* idx, const lte_c32* symbols, int32_t T,
                    lte_c32* tx, lte_c32* qam_out, double* stats, int32_t B, int32_t S,
                    void* stream);

/* --- PAPR / CCDF engine ------------------------------------------------------
 * lte_tx_papr: lte_tx_map_ifft with a per-OFDM-symbol PAPR epilogue over the N useful samples
 * (no CP), replacing OFDMSystem.calculate_papr_without_cp (core/ofdm_system.py:173-229) and the
 * collection loop of collect_papr_for_all_modulations (:648-735).  tx may be NULL: the
 * time-domain stream is then never written (sweep mode).  Outputs, each optional (at least one):
 * papr_db [B*T][S] = 10 log10(peak/mean) (0 for an all-zero symbol); peak_mean [B*T][S][2]
 * floats; hist [hist_bins] counts of papr_db in bins of width hist_step from hist_lo (values
 * outside fall into the first / last bin), accumulated with atomics, caller zeroes.
 * lte_papr_symbols: the same outputs for an existing stream x [rows][S*L]; include_cp != 0
 * takes all L samples of a symbol (OFDMSystem.calculate_papr_per_symbol, :116-171).
 * lte_histogram: histogram of any float array with the same bin rule. */
int lte_tx_papr(const lte_plan*, const uint8_t* idx, const lte_c32* symbols, int32_t T, lte_c32* tx,
                double* stats, float* papr_db, float* peak_mean, unsigned long long* hist,
                float hist_lo, float hist_step, int32_t hist_bins, int32_t B, int32_t S, void* stream);
int lte_papr_symbols(const lte_plan*, const lte_c32* x, int32_t include_cp, float* papr_db,
                     float* peak_mean, unsigned long long* hist, float hist_lo, float hist_step,
                     int32_t hist_bins, int64_t rows, int32_t S, void* stream);
int lte_histogram(const float* x, int64_t n, float lo, float step, int32_t bins,
                  unsigned long long* hist, void* stream);

/* --- SC-FDM M-point unitary DFT / IDFT ------------------------------------------------
 * replaces DFTPrecodifier.precoding / IDFTDecodifier.decoding and the SC_FDMPrecodifier /
 * SC_FDMDecodifier wrappers (core/dft_precoding.py:67-93, :188-216, :254-348):
 * out[r][k] = sum_n in[r][n] exp(-+ j 2 pi k n / M) / sqrt(M) for every row of M symbols
 * (inverse != 0 selects the + sign).  Any M <= 1024 (Bluestein chirp-z on the power-of-two
 * FFT core; the LTE data-subcarrier counts 62..999 are not 2-3-5 smooth).  The chirp tables of an M are
 * plan state: lte_plan_add_dft(plan, M) builds them once (synchronous, like lte_plan_create); lte_dft_m on an
 * M the plan does not hold returns LTE_ERR_INVALID_ARG. */
int lte_plan_add_dft(lte_plan* plan, int32_t M);
/* lte_dft_qam = lte_qam_map + lte_dft_m(forward) in one kernel (the QAM map rides in the transform's load; same
 * arithmetic): idx [rows][M] -> out [rows][M]. */
int lte_dft_qam(const lte_plan*, const uint8_t* idx, lte_c32* out, int32_t M, int64_t rows, void* stream);
int lte_dft_m(const lte_plan*, const lte_c32* in, lte_c32* out, int32_t M, int32_t inverse,
              int64_t rows, void* stream);

/* --- stage 3 channel: time-domain tapped delay line with per-sample Jakes fading
 * replaces RayleighChannel.jakes_fading/.filter (core/rayleighchannel.py:20-58) for
 * R receive antennas and T transmit antennas (OFDMChannel.transmit_simo,
 * core/ofdm_core.py:361-412).  tx: [B][T][n]; phases: [B][R][T][taps][16] uniform
 * draws u in [0,1) (phi = 2*pi*u); faded: [B][R][n], the sum over TX;
 * power: [B][R] sum of |faded|^2 (double, accumulated with atomics, caller zeroes).
 * With num_taps == 0 (channel_type 'awgn') the link is the identity: only `power`
 * (= sum |tx|^2 for every antenna) is produced and `faded` is not written.
 * workspace: lte_channel_tdl_workspace_bytes() bytes (Jakes polynomial coefficients; may be NULL
 * for the identity link).  Accuracy: the fading process is a Taylor polynomial per block of samples
 * (remainder < 2e-8 of |h|) stepped linearly inside each thread's 8 samples, which adds
 * (2 pi fD/fs 3.5)^2 / 2 -- 3e-6 for Vehicular_B at 120 km/h and 1.92 MHz; configurations where this
 * term exceeds 1e-4 (fD/fs > 6.4e-4) return LTE_ERR_UNSUPPORTED. */
int64_t lte_channel_tdl_workspace_bytes(const lte_plan*, const lte_channel_desc* ch, int32_t B,
                                        int32_t R, int32_t T, int64_t n);   /* bytes, or LTE_ERR_* (< 0) */
int lte_channel_tdl(const lte_plan*, const lte_channel_desc* ch, const lte_c32* tx,
                    const float* phases, lte_c32* faded, double* power, void* workspace, int32_t B,
                    int32_t R, int32_t T, int64_t n, void* stream);

/* AWGN of AWGNChannel.transmit / RayleighMultiPathChannel.transmit
 * (core/channel.py:46-66, :216-232): sigma = sqrt(power/n / snr_lin / 2) per row.
 * z: [rows][n] unit normals (re, im) to replay the reference's draws, or NULL to
 * draw them from Philox2x32-10 keyed (seed, row_id0 + row, sample).
 * Output row i reads input row i / x_div, so R antennas can share one TX stream
 * (x_div = R on the AWGN channel type, 1 otherwise); power: [rows]; snr_lin: [rows]. */
int lte_awgn_add(const lte_plan*, const lte_c32* x, int32_t x_div, const double* power,
                 const float* snr_lin, const lte_c32* z, uint64_t seed, uint64_t row_id0,
                 lte_c32* y, int64_t rows, int64_t n, void* stream);

/* --- stage 2 RX: CP strip + FFT/sqrt(N) -------------------------------------------
 * replaces LTEReceiver._demodulate_ofdm_stream (core/lte_receiver.py:444-491).
 * rx: [rows / rx_div][S*L] (row i reads input row i / rx_div); Y: [rows][S][nk] for
 * the chosen window.
 * Optional fused noise (engine mode): if power != NULL the kernel adds AWGN with the
 * sigma rule of lte_awgn_add.  noise_domain 0: sigma*(z_re + j z_im) per time sample
 * before the FFT, from the injected normals z or (z == NULL) from Philox with the very
 * keying of lte_awgn_add, so fused and unfused paths agree bit for bit.  noise_domain 1
 * (z must be NULL): the same-variance white Gaussian noise is drawn directly on the
 * output bins after the unitary FFT -- identical in distribution, fewer draws. */
int lte_rx_fft(const lte_plan*, const lte_c32* rx, int32_t rx_div, const double* power,
               const float* snr_lin, const lte_c32* z, int32_t noise_domain, uint64_t seed,
               uint64_t row_id0, lte_c32* Y, int window,
               int64_t rows, int32_t S, void* stream);

/* --- stage 4: CRS least squares + linear interpolation, one estimate per 14 symbols
 * replaces LTEChannelEstimator.estimate_channel/_interpolate_channel
 * (core/lte_receiver.py:40-133) and LTEReceiver._estimate_channel_periodic (:360-411).
 * Y: [rows][S][nk]; H: [rows][ceil(S/14)][nk]; pilot_set selects the TX pilot table. */
int lte_crs_ls_interp(const lte_plan*, const lte_c32* Y, lte_c32* H, int window, int pilot_set,
                      int64_t rows, int32_t S, void* stream);

/* --- stage 5 equalisers --------------------------------------------------------------
 * ZF: Y/(H+1e-6) on data bins (LTEEqualizerZF.equalize, core/lte_receiver.py:154-180);
 * H == NULL skips the division (enable_equalization=False, core/lte_receiver.py:294-300).
 * MRC over R antennas: sum conj(H_r) Y_r / (sum |H_r|^2 + 1e-10)
 * (OFDMSimulator._combine_symbols_mrc, core/ofdm_core.py:1405-1534).
 * Y: [B][R][S][nk]; H: [B][R][ceil(S/14)][nk]; out: [B][S][Nd] data symbols. */
int lte_equalize_zf(const lte_plan*, const lte_c32* Y, const lte_c32* H, lte_c32* out, int window,
                    int64_t B, int32_t S, void* stream);
int lte_equalize_mrc(const lte_plan*, const lte_c32* Y, const lte_c32* H, lte_c32* out, int window,
                     int64_t B, int32_t R, int32_t S, void* stream);

/* --- fused stage 3 + CP strip + FFT (sweep engine) -----------------------------------------------
 * lte_channel_tdl followed by lte_rx_fft(noise_domain none) in ONE kernel for T = 1: the faded
 * time-domain streams [B][R][S*L] are formed in registers / shared memory and never written.
 * replaces RayleighChannel.jakes_fading/.filter (core/rayleighchannel.py:20-58) for the R links
 * of OFDMChannel.transmit_simo (core/ofdm_core.py:361-412) and
 * LTEReceiver._demodulate_ofdm_stream (core/lte_receiver.py:444-491).
 * tx: [B][S*L]; phases as in lte_channel_tdl with T = 1; Y: [B*R][S][nk] noise-free;
 * power: [B][R] sum |faded|^2 over the whole stream including the CP samples, accumulated with
 * atomics, caller zeroes.  The AWGN is then added lazily by the *_awgn consumers below.
 * Returns LTE_ERR_UNSUPPORTED for the identity link (num_taps = 0) and when the Doppler spread
 * is too large for one Jakes polynomial (degree <= 6) per OFDM symbol, pi fD L / fs > 0.25 -- about
 * 1.1 kHz at 20 MHz, 1.7 kHz at 1.25 MHz: use the staged pair. */
int64_t lte_channel_rx_fft_workspace_bytes(const lte_plan*, const lte_channel_desc* ch, int32_t B,
                                           int32_t R, int32_t S);             /* bytes, or LTE_ERR_* (< 0) */
int lte_channel_rx_fft(const lte_plan*, const lte_channel_desc* ch, const lte_c32* tx,
                       const float* phases, lte_c32* Y, double* power, void* workspace, int window,
                       int32_t B, int32_t R, int32_t S, void* stream);
/* The same for T > 1 transmit antennas (SFBC, spatial multiplexing): the R x T independently faded links of
 * ChannelSimulator.transmit_spatial_multiplexing / the MIMO branch of OFDMChannel (core/channel.py:399-466,
 * core/ofdm_core.py:1850-2258) summed per receive antenna, stream power, CP strip and FFT in one kernel.
 * tx: [B][T][S*L]; phases: [B][R][T][taps][16] as in lte_channel_tdl; Y: [B*R][S][nk] noise-free; power [B][R].
 * The AWGN joins in the consumers: lte_crs_ls_interp_awgn, lte_sfbc_decode_count(awgn), lte_mimo_detect(awgn). */
int64_t lte_channel_rx_fft_mimo_workspace_bytes(const lte_plan*, const lte_channel_desc* ch, int32_t B,
                                                int32_t R, int32_t T, int32_t S);
int lte_channel_rx_fft_mimo(const lte_plan*, const lte_channel_desc* ch, const lte_c32* tx,
                            const float* phases, lte_c32* Y, double* power, void* workspace, int window,
                            int32_t B, int32_t R, int32_t T, int32_t S, void* stream);

/* --- spectral fading link (sweep engine, T = 1, low Doppler) -------------------------------------
 * The same reference stages as lte_tx_map_ifft + lte_channel_rx_fft -- QAMModulator.bits_to_symbols,
 * ResourceMapper.map_symbols, OFDMModulator ifft (core/modulator.py:61-88,214-302,
 * core/resource_mapper.py:181-223), RayleighChannel.jakes_fading/.filter for the R links
 * (core/rayleighchannel.py:20-58, core/ofdm_core.py:361-412), the stream power behind the AWGN
 * (core/channel.py:216-218) and LTEReceiver._demodulate_ofdm_stream (core/lte_receiver.py:444-491) --
 * without a time-domain stream in HBM and with one forward transform per OFDM symbol instead of one
 * per antenna.  Valid when every Jakes process is linear over one OFDM symbol to 5e-7
 * (pi fD L / fs <= 1.41e-3, i.e. 3 km/h at 2 GHz) and every tap delay fits in the cyclic prefix:
 *   Y_r[k] = sum_t exp(-2 pi j k d_t / N) { (c0 + d_t c1) X[k] + c1 (G[k] - N T_t[k]) }
 * (csrc/spectral.cu).  lte_tx_spectral: idx [B][S][Nd] -> G [B*S][Nc] = fft((n - n_c) u[n]) / sqrt(N)
 * on the occupied window and tail [B*S][cp] = the last cp samples of every symbol.
 * lte_channel_spectral: idx, G, tail, phases [B][R][taps][16] -> Y [B*R][S][Nc] (LTE_WINDOW_USEFUL,
 * noise-free) and power [B][R] (accumulated, caller zeroes) exactly as lte_channel_rx_fft defines
 * them.  With Ypilot != NULL the grid leaves in the COMPACT layout of the sweep consumers
 * (lte_crs_ls_compact, lte_mrc_demap_count_compact, described there): Y = Ydata holds the data bins
 * only and Ypilot the pilot bins of every slot's first symbol -- the only pilots
 * LTEReceiver._estimate_channel_periodic (core/lte_receiver.py:360-411) reads.  workspace: lte_channel_spectral_workspace_bytes() bytes of device memory owned by the caller
 * (Jakes coefficients; contents are scratch).  Both the size query and the launcher return
 * LTE_ERR_UNSUPPORTED outside the validity range: use lte_tx_map_ifft + lte_channel_rx_fft. */
int lte_tx_spectral(const lte_plan*, const uint8_t* idx, lte_c32* G, lte_c32* tail, int32_t B, int32_t S,
                    void* stream);
int64_t lte_channel_spectral_workspace_bytes(const lte_plan*, const lte_channel_desc* ch, int32_t B,
                                             int32_t R, int32_t S);
int lte_channel_spectral(const lte_plan*, const lte_channel_desc* ch, const uint8_t* idx, const lte_c32* G,
                         const lte_c32* tail, const float* phases, lte_c32* Y, lte_c32* Ypilot, double* power,
                         void* workspace, int32_t B, int32_t R, int32_t S, void* stream);

/* --- lazy frequency-domain AWGN for the sweep engine -------------------------------------------
 * The noise lte_rx_fft(noise_domain = 1) would add to grid element (row, symbol, bin) is a pure
 * function of (seed, row_id0 + row, symbol, bin).  These variants take a noise-free Y and add that
 * noise while reading it (bit-identical result), so the noisy grid never exists in HBM and only
 * the elements that are consumed get a noise sample.  power [rows] as produced by
 * lte_channel_tdl / lte_channel_rx_fft, snr_lin [rows]; rows = B*R. */
typedef struct {
    const double* power;
    const float* snr_lin;
    uint64_t seed;
    uint64_t row_id0;
    /* lte_mrc_demap_count_awgn only.  0: one draw per (antenna, symbol, bin), the very draws of the
     * RX epilogue.  1: one draw per combiner output, scaled by sqrt(sum_r |H_r|^2 sigma_r^2) -- the
     * exact distribution of sum_r conj(H_r) sigma_r w_r given H, at 1/R of the generator work
     * (different sample values, same BER statistics). */
    int32_t combine;
} lte_awgn_desc;
int lte_crs_ls_interp_awgn(const lte_plan*, const lte_c32* Y, lte_c32* H, int window, int pilot_set,
                           int64_t rows, int32_t S, const lte_awgn_desc* awgn, void* stream);
int lte_mrc_demap_count_awgn(const lte_plan*, const lte_c32* Y, const lte_c32* H, const uint8_t* idx_tx,
                             unsigned long long* errors, int window, int64_t nbits, int64_t B, int32_t R,
                             int32_t S, const lte_awgn_desc* awgn, void* stream);
/* lte_equalize_zf on a noise-free grid (SISO through lte_channel_rx_fft): the AWGN joins as the data bins are read. */
int lte_equalize_zf_awgn(const lte_plan*, const lte_c32* Y, const lte_c32* H, lte_c32* out, int window,
                         int64_t B, int32_t S, const lte_awgn_desc* awgn, void* stream);

/* Compact sweep layout (produced by lte_channel_spectral with Ypilot != NULL): the grid travels as
 *   Ydata  [B*R][S][2 ndp]            data symbol d of the OFDM symbol at element d, ndp = ceil(Nd / 2)
 *                                     (odd Nd: the last element of a row is padding)
 *   Ypilot [B*R][ceil(S/14)][2 npp]   pilot i of every slot's first symbol at element i, npp = ceil(Np / 2)
 * i.e. exactly the elements LTEReceiver._estimate_channel_periodic (core/lte_receiver.py:360-411) and
 * _combine_symbols_mrc (core/ofdm_core.py:1484-1532) read, in rows of 16-byte pairs.  lte_plan_compact_shape
 * gives ndp and npp.
 * lte_crs_ls_compact: the LS step of LTEChannelEstimator.estimate_channel (core/lte_receiver.py:62-87),
 * Hp [B*R][ceil(S/14)][Np] = (Ypilot [+ AWGN]) / pilot; lte_mrc_demap_count_compact interpolates
 * between the two pilots around each data bin (core/lte_receiver.py:98-133: edge hold + np.linspace, the
 * very operations of lte_crs_ls_interp), combines, slices and counts.  awgn may be NULL (noise-free
 * grid); with awgn the draws are those of lte_crs_ls_interp_awgn / lte_mrc_demap_count_awgn, so the
 * error counts are bit-identical to the windowed layout.  Single pilot set (T = 1) only. */
int lte_plan_compact_shape(const lte_plan*, int32_t* ndp, int32_t* npp);
int lte_crs_ls_compact(const lte_plan*, const lte_c32* Ypilot, lte_c32* Hp, int64_t rows, int32_t S,
                       const lte_awgn_desc* awgn, void* stream);
int lte_mrc_demap_count_compact(const lte_plan*, const lte_c32* Ydata, const lte_c32* Hp,
                                const uint8_t* idx_tx, unsigned long long* errors, int64_t nbits, int64_t B,
                                int32_t R, int32_t S, const lte_awgn_desc* awgn, void* stream);
/* lte_crs_ls_compact + lte_mrc_demap_count_compact in one launch (the sweep's form): every CTA forms the LS
 * estimates of the <= 28 pilots its 128 data bins interpolate between from Ypilot itself, with the very
 * operations and lazy-AWGN draws of lte_crs_ls_compact, so the counts are bit-identical to the two calls and
 * no estimate tensor exists (LTEChannelEstimator.estimate_channel + _interpolate_channel,
 * core/lte_receiver.py:62-133, inside the combiner of core/ofdm_core.py:1484-1532). */
int lte_crs_mrc_demap_count_compact(const lte_plan*, const lte_c32* Ydata, const lte_c32* Ypilot,
                                    const uint8_t* idx_tx, unsigned long long* errors, int64_t nbits, int64_t B,
                                    int32_t R, int32_t S, const lte_awgn_desc* awgn, void* stream);

/* --- SFBC Alamouti transmit diversity (2 TX) ---------------------------------------------
 * lte_sfbc_encode replaces SFBCAlamouti.encode (core/sfbc_alamouti.py:45-78) fused with the QAM
 * map: idx [B][S][2*(Nd/2)] symbol indices (or `symbols`, complex, same shape) ->
 * out [B][2][S][Nd] per-antenna data symbols ready for lte_tx_map_ifft(symbols=out, T=2) on a
 * plan with two pilot sets (SFBCResourceMapper.map_sfbc_to_grid, :212-264); with an odd Nd the
 * last data bin is nulled (:196-200).  qam_out (optional): [B][S][2*(Nd/2)] mapped symbols.
 * lte_sfbc_decode replaces SFBCAlamouti.decode (:80-163) for every RX antenna followed by the
 * arithmetic mean over antennas (core/ofdm_core.py:2204): Y [B][R][S][nk]; H0/H1
 * [B*R][ceil(S/14)][nk] (lte_crs_ls_interp with pilot_set 0 / 1); out [B][S][2*(Nd/2)]. */
int lte_sfbc_encode(const lte_plan*, const uint8_t* idx, const lte_c32* symbols, lte_c32* out,
                    lte_c32* qam_out, int64_t B, int32_t S, void* stream);
int lte_sfbc_decode(const lte_plan*, const lte_c32* Y, const lte_c32* H0, const lte_c32* H1,
                    lte_c32* out, int window, int64_t B, int32_t R, int32_t S, void* stream);
/* lte_sfbc_encode fused into lte_tx_map_ifft(T = 2): idx [B][S][2*(Nd/2)] -> tx [B][2][S*L]; the encoded
 * symbols never exist in memory (antenna 0 reads idx[c], antenna 1 idx[c ^ 1], conjugation / negation by sign).
 * Same samples as lte_sfbc_encode + lte_tx_map_ifft(symbols). */
int lte_tx_sfbc_ifft(const lte_plan*, const uint8_t* idx, lte_c32* tx, int32_t B, int32_t S, void* stream);
/* lte_sfbc_decode fused with lte_demap_count (QAMModulator.symbols_to_bits + calculate_ber,
 * core/modulator.py:90-112, core/ofdm_core.py:245-268): the decoded symbols are sliced and compared with
 * idx_tx [B][S][2*(Nd/2)] in registers; errors[b] (caller zeroes) += bit errors among the first nbits bits
 * of stream b.  Same decoded values, hence the same counts, as the two separate calls.
 * awgn (optional): Y is noise free (lte_channel_rx_fft_mimo) and the noise lte_rx_fft(noise_domain = 1)
 * would have added joins while the bins are read; H0 / H1 then come from lte_crs_ls_interp_awgn. */
int lte_sfbc_decode_count(const lte_plan*, const lte_c32* Y, const lte_c32* H0, const lte_c32* H1,
                          const uint8_t* idx_tx, unsigned long long* errors, int64_t nbits, int window,
                          int64_t B, int32_t R, int32_t S, const lte_awgn_desc* awgn, void* stream);

/* --- spatial multiplexing (TM4-like, up to 8 TX / 8 RX / 4 layers) ------------------------
 * lte_sm_precode replaces LayerMapper.map_to_layers (core/layer_mapper.py:35-86) and the
 * per-subcarrier x_k = W layers[:, k] loop of simulate_spatial_multiplexing
 * (core/ofdm_core.py:2611-2640): idx [B][S][Nd] (or complex `symbols`) -> out [B][T][S][Nd];
 * only the first ceil(Nd / L) data bins carry data, as in the reference.  W_host: [T][L]
 * precoder (host memory, row major).
 * lte_flat_mimo replaces the flat branch of ChannelSimulator.transmit_spatial_multiplexing
 * (core/channel.py:467-480): out[b][r] = sum_t h[b][r][t] tx[b][t]; power [B][R] accumulates
 * sum |out|^2 (caller zeroes).  The multipath branch is lte_channel_tdl with T > 1.
 * lte_mimo_detect replaces MIMODetector.detect (core/mimo_detector.py:55-369) on
 * H_eff = H W per data position: Y [B][R][S][nk]; H [T][B*R][S][nk] (per-symbol estimates, one
 * lte_crs_ls_interp per TX pilot set); out [B][S][Nd] in the LayerMapper.demap_from_layers order
 * (core/layer_mapper.py:88-115).  detector: LTE_DET_*; SIC slices with the plan's constellation.
 * H may be NULL: the detector then forms the per-symbol CRS estimate of every TX antenna at its bin itself
 * from Y's pilot bins (LS at the two neighbouring pilots of set t, linear in between, edge hold) -- bit for bit
 * the value lte_crs_ls_interp(pilot_set = t, rows = B*R*S, S = 1) would have written, without the T passes
 * over Y and the [T][B*R][S][nk] tensor (the plan must carry the T pilot sets).  Hpilot (optional, with H = NULL):
 * the LS estimates at the pilots from lte_crs_ls_pilots; the detector then only interpolates (with lazy AWGN
 * every pilot gets its noise sample once instead of once per data bin next to it).
 * awgn (optional): Y is noise free and the AWGN joins while data and pilot bins are read (lazy AWGN, as in
 * lte_mrc_demap_count_awgn); with a non-NULL H the estimates must already include it (lte_crs_ls_interp_awgn).
 * sigma2_streams (optional, device, double [B]): one noise variance per stream, which lets a sweep put
 * all its SNR points into one launch; NULL = the scalar `sigma2` for every stream. */
int lte_sm_precode(const lte_plan*, const uint8_t* idx, const lte_c32* symbols, const lte_c32* W_host,
                   int32_t T, int32_t L, lte_c32* out, lte_c32* qam_out, int64_t B, int32_t S,
                   void* stream);
int lte_flat_mimo(const lte_plan*, const lte_c32* tx, const lte_c32* h, lte_c32* out, double* power,
                  int64_t B, int32_t R, int32_t T, int64_t n, void* stream);
/* lte_crs_ls_pilots: the LS step of the per-symbol estimate alone, for every pilot set of the plan in one launch
 * (MIMOChannelEstimator / estimate_channel per TX antenna, core/lte_receiver.py:62-87): Y [rows][S][nk] ->
 * Hp [nsets][rows*S][Np] (the first pilots-of-set-t entries of a row are used); awgn optional (lazy AWGN). */
int lte_crs_ls_pilots(const lte_plan*, const lte_c32* Y, lte_c32* Hp, int window, int64_t rows, int32_t S,
                      const lte_awgn_desc* awgn, void* stream);
int lte_mimo_detect(const lte_plan*, const lte_c32* Y, const lte_c32* H, const lte_c32* Hpilot, const lte_c32* W_host,
                    int32_t T, int32_t L, double sigma2, const double* sigma2_streams, int32_t detector,
                    lte_c32* out, int window, int64_t B, int32_t R, int32_t S, const lte_awgn_desc* awgn,
                    void* stream);

/* --- beamforming (rank-1 precoding over a flat R x T channel) ---------------
 * The reference path OFDMSimulator.simulate_beamforming (core/ofdm_core.py:2260-2477) stays in
 * the frequency domain: per OFDM symbol  x = W s,  y = H x + n,  MRC with H_eff = H W.
 * lte_random_channel: h [B][R][T] ~ CN(0,1), the (randn + j randn)/sqrt(2) draw of :2347-2348,
 * Philox keyed (seed, stream_id0 + b).
 * lte_bf_weights replaces CSIFeedback.generate_feedback -> LTECodebook.select_best_pmi
 * (core/csi_feedback.py:166-196, core/codebook_lte.py:332-373) and
 * BeamformingPrecoder.calculate_mrt_weights / calculate_beamforming_gain
 * (core/beamforming_precoder.py:41-66, :176-201) for B channel matrices at once:
 * codebook_host [ncb][T] rank-1 precoders (host memory); mode LTE_BF_MRT | LTE_BF_CODEBOOK selects
 * which precoder is written to W [B][T]; heff [B][R] = H W; pmi [B] (optional) is always the
 * codebook argmax, first maximum wins; gain_db [B] (optional).
 * lte_bf_link replaces the per-symbol loop, the MRC and the demap/BER of :2359-2439:
 * idx [B][S][Nd]; noise_std [B] = sqrt(10^(-snr/10) / 2); z (optional) [B][S][2][R][Nd] replayed
 * unit normals (real block, then imaginary block, as the reference draws them), else Philox keyed
 * (seed, row_id0 + b*R + r, s*Nd + d); out (optional) [B][S][Nd] equalised symbols; errors
 * (optional) [B] uint64 accumulated over the first nbits bits of each stream (caller zeroes). */
int lte_random_channel(lte_c32* h, int64_t B, int32_t R, int32_t T, uint64_t seed, uint64_t stream_id0,
                       void* stream);
int lte_bf_weights(const lte_c32* h, const lte_c32* codebook_host, int32_t ncb, int32_t mode, lte_c32* W,
                   lte_c32* heff, int32_t* pmi, float* gain_db, int64_t B, int32_t R, int32_t T,
                   void* stream);
/* lte_rank_feedback replaces RankAdaptation.get_feedback with its default methods -- RI from the eigenvalues of
 * H^H H with the SNR gates (core/rank_adaptation.py:41-130), PMI by the capacity metric over the rank's
 * codebook, first maximum wins (:148-211) -- for n channel matrices at once, in fp64 like the reference:
 * H [n][R][T] (R <= 8, T <= 4); snr_db [n] (device); codebook [max_rank][ncb_stride][T][4] (device; entry
 * (rank - 1, pmi) holds W [T][rank] in its first `rank` columns); ncb_host [max_rank] codebook sizes per rank
 * (host); ri / pmi [n] int32 (device).  The spatial-multiplexing sweep uses it for the reference's
 * H_initial feedback of simulate_spatial_multiplexing (core/ofdm_core.py:2573-2583), one matrix per
 * (SNR point, feedback block), in one launch. */
int lte_rank_feedback(const lte_c32* H, const double* snr_db, const lte_c32* codebook, int32_t ncb_stride,
                      const int32_t* ncb_host, double rank_threshold, int32_t max_rank, int32_t* ri, int32_t* pmi,
                      int64_t n, int32_t R, int32_t T, void* stream);
int lte_bf_link(const lte_plan*, const uint8_t* idx, const lte_c32* h, const lte_c32* W,
                const lte_c32* heff, const float* noise_std, const float* z, uint64_t seed,
                uint64_t row_id0, lte_c32* out, unsigned long long* errors, int64_t nbits, int64_t B,
                int32_t R, int32_t T, int32_t S, void* stream);

/* --- coded chain (CRC + turbo code + soft demapping) ----------------------
 * Replaces the channel-coding layers OFDMSimulator.simulate_siso_coded adds around the SISO chain
 * (core/ofdm_core.py:925-1338).  One stream = one transport block of A bits.  The block structure
 * depends only on A; the host lays it out once (lte_b200/coding.py) as device-resident int32 tables:
 *   blk [C][LTE_BLK_COLS] = {K, filler F, info bits n, offset into (bits ++ crc24a), offset into the
 *        code-block row, offset into the encoded row, has CRC-24B, offset into pi_tab}
 *   rm_table [sumE], dm_table [sumE] (sumE = sum_r 3 K_r + 12; -1 = constant 0), pi_tab (QPP
 *   permutations pi(i) = (f1 i + f2 i^2) mod K of the distinct K).
 * lte_tb_encode: attach_crc24a (core/channel_coding/crc.py:204-232), segment_code_blocks
 *   (segmentation.py:66-199), turbo_encode (turbo_encoder.py:112-259, whose RSC emits the feedback
 *   bit as "systematic"), rate_match_turbo with E = 3K+12, rv 0 (rate_matching.py:27-229).
 *   bits [B][A] bytes 0/1 -> coded [B][sumE]; crc [B][24], cb [B][sumK], enc [B][sumE] are scratch.
 * lte_symbol_interleave: QAM map + the rows x Nd block interleaver of core/ofdm_core.py:1037-1060
 *   (complex-zero padding): idx [B][nsym] -> out [B][rows*Nd], ready for lte_tx_map_ifft(symbols=out).
 * lte_soft_demap: de-interleave (:1180-1215), per-symbol noise variance (:1228-1250) and the LLRs of
 *   _calculate_llrs_qpsk/_16qam/_64qam (:791-923): data [B][rows*Nd] equalised symbols in received
 *   order, H [B][ceil(rows/14)][nk] channel estimates (window as in lte_crs_ls_interp), sigma2 [B] =
 *   10^(-snr/10), fading != 0 selects the |H|^2-scaled variance -> llr [B][nsym*bits_per_symbol].
 * lte_tb_decode: rate_dematching_turbo (rate_matching.py:297-396), turbo_decode in its default
 *   max-log mode (turbo_decoder.py:158-446; `iterations` full iterations + the final decoder-1
 *   pass; logmap != 0 selects the exact max* = log(e^a + e^b) of set_decoder_mode(False), :35-120), desegment_code_blocks (segmentation.py:202-270), check_crc24a and the BER count
 *   (core/ofdm_core.py:1283-1307).  llr [B][sumE]; dematched [B][sumE] and work
 *   [ceil(B*C/4)*4][lte_tb_decode_work_floats(Kmax)] floats and cbdec [B][sumK] are scratch; bits_tx (optional)
 *   [B][A]; outputs (each optional) bits_rx [B][A], crc_ok [B] int32, errors [B] uint64. */
int lte_tb_encode(const uint8_t* bits, int64_t A, const int32_t* blk, int32_t C, int64_t sumK,
                  int64_t sumE, const int32_t* rm_table, const int32_t* pi_tab, uint8_t* crc,
                  uint8_t* cb, uint8_t* enc, uint8_t* coded, int64_t B, void* stream);
int lte_symbol_interleave(const lte_plan*, const uint8_t* idx, int64_t nsym, int32_t rows, lte_c32* out,
                          int64_t B, void* stream);
int lte_soft_demap(const lte_plan*, const lte_c32* data, const lte_c32* H, int window,
                   const float* sigma2, int32_t fading, int64_t nsym, int32_t rows, float* llr, int64_t B,
                   void* stream);
int64_t lte_tb_decode_work_floats(int32_t Kmax);
int lte_tb_decode(const float* llr, const int32_t* blk, int32_t C, int64_t sumK, int64_t sumE, int32_t Kmax,
                  const int32_t* dm_table, const int32_t* pi_tab, int32_t iterations, int32_t logmap, float* dematched,
                  float* work, uint8_t* cbdec, int64_t A, const uint8_t* bits_tx, uint8_t* bits_rx,
                  int32_t* crc_ok, unsigned long long* errors, int64_t B, void* stream);

/* Stage-level entry points behind the reference's module functions (core/channel_coding/__init__.py):
 * lte_crc_bits -> calculate_crc24a/24b/16 (crc.py:89-209; poly without its leading term, e.g. 0x864CFB);
 * lte_turbo_encode_blocks -> turbo_encode (turbo_encoder.py:214-313) on ready code blocks cb [B][sumK];
 * lte_turbo_decode_blocks -> turbo_decode (turbo_decoder.py:338-446) on decoder-order LLRs dl [B][sumE];
 *   with iterations = 0 it is one LogMAPDecoder.decode pass (:158-293): apriori (optional) [B][sumK] a-priori
 *   LLRs of the K data steps, apost (optional) [B][sumK + 3C] a-posteriori LLRs of all K+3 steps per block;
 * lte_gather_u8 / lte_gather_f32 -> qpp_interleave / qpp_deinterleave, sub_block_interleaver /
 *   _deinterleaver, rate_match_turbo / rate_dematching_turbo as index tables (-1 = constant 0). */
int lte_crc_bits(const uint8_t* bits, int64_t n, uint32_t poly, int32_t len, uint8_t* out, int64_t B,
                 void* stream);
int lte_turbo_encode_blocks(const uint8_t* cb, const int32_t* blk, int32_t C, int64_t sumK, int64_t sumE,
                            const int32_t* pi_tab, uint8_t* enc, int64_t B, void* stream);
int lte_turbo_decode_blocks(const float* dl, const int32_t* blk, int32_t C, int64_t sumK, int64_t sumE,
                            int32_t Kmax, const int32_t* pi_tab, int32_t iterations, int32_t logmap,
                            float* work, uint8_t* cbdec, const float* apriori, float* apost, int64_t B,
                            void* stream);
int lte_gather_u8(const uint8_t* src, int64_t n_src, const int32_t* table, int64_t n, uint8_t* out, int64_t B,
                  void* stream);
int lte_gather_f32(const float* src, int64_t n_src, const int32_t* table, int64_t n, float* out, int64_t B,
                   void* stream);

/* --- stage 6: hard demap + bit-error count -------------------------------------------
 * replaces QAMModulator.symbols_to_bits (core/modulator.py:90-112) and
 * OFDMReceiver.calculate_ber (core/ofdm_core.py:245-268).  syms: [B][nsym];
 * idx_tx: [B][nsym] transmitted indices (may be NULL: demap only); idx_rx (optional):
 * [B][nsym] decided indices; errors: [B] uint64, atomically accumulated (caller zeroes);
 * only the first nbits bits of each stream are compared (core/ofdm_core.py:712-718). */
int lte_demap_count(const lte_plan*, const lte_c32* syms, const uint8_t* idx_tx, uint8_t* idx_rx,
                    unsigned long long* errors, int64_t nsym, int64_t nbits, int64_t B,
                    void* stream);

/* Fused stage 5+6 for the sweep engine: MRC -> slicer -> XOR/popcount without
 * writing the combined symbols. */
int lte_mrc_demap_count(const lte_plan*, const lte_c32* Y, const lte_c32* H, const uint8_t* idx_tx,
                        unsigned long long* errors, int window, int64_t nbits, int64_t B, int32_t R,
                        int32_t S, void* stream);

/* Measurement helper (bench.py's compute roofline): launches a register-only loop of independent packed
 * fp32 multiply-adds (fma.rn.f32x2, the instruction the FMA-bound kernels of this library issue) on the current
 * device and returns the flops of that launch (< 0: LTE_ERR_*); the caller times the launch with CUDA events.
 * sink: (SM count * 8 * 256) floats of scratch. */
int64_t lte_fp32_peak_launch(float* sink, int32_t iters, void* stream);

/* Engine helper: Philox-generated uniform symbol indices [B][nsym] keyed (seed, stream id). */
int lte_random_indices(const lte_plan*, uint8_t* idx, int64_t nsym, int64_t B, uint64_t seed,
                       uint64_t stream_id0, void* stream);
/* Engine helper: Philox-generated Jakes phase draws u in [0,1): [B][links][taps][16]. */
int lte_random_phases(float* phases, int64_t count_per_stream, int64_t B, uint64_t seed,
                      uint64_t stream_id0, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LTE_B200_H */
