// march_plan.inl -- host-side plan of the exact-integer separable kernels (vpp_march_kernels.cu).
// Included by tsvpp_capi.cpp inside its anonymous namespace (it uses Geometry, AxisSample, bicubic_taps,
// DeviceArena, round_up, env_int from there).
//
// Everything here is about ONE question: is the reference's arithmetic for this geometry exact in small
// integers?  If yes the tables below are built; if any check fails the caller keeps the float plan.

struct MarchInputs {
    Geometry g;
    int rs = 0, ok = 0;
    bool need_uv = true;
    int W = 0, H = 0, pitch_y = 0, pitch_uv = 0;
    int dw = 0, dh = 0, cw2 = 0, ch2 = 0;
    int swap_rb = 0;
    const std::vector<AxisSample> *xs = nullptr, *ys = nullptr;
    const std::vector<int> *tap_l = nullptr, *tap_c = nullptr, *lo_c = nullptr;  // [dw][4] absolute byte positions
    const std::vector<int> *lo_l = nullptr, *hi_l = nullptr, *hi_c = nullptr;
    const std::vector<double> *cxw = nullptr, *cyw = nullptr;  // bicubic: [dw][4], [>= dh][4] FP64 weights (device pow)
    int K = 1, KX = 1, PX = 1, PY = 1;
    const std::vector<float> *patx = nullptr, *paty = nullptr;  // area patterns [PX][KX], [PY][K]
    size_t smem_cap = 0;
};

struct MarchPlanHost {
    MarchPlanDev dev{};
    void *dev_tables = nullptr;
    size_t smem = 0;
    int threads = 0;
    int ctas_per_sm = 1;
    long fetch_bytes_per_frame = 0;
    std::string desc;
};

// smallest q <= qmax with v * 2^q within `tol` (absolute, in units of v) of an integer for every v
// (all |v| <= 4); -1 if none.  tol = 0 asks for exactly dyadic values.
static int dyadic_q(const double *v, size_t n, int qmax, double tol = 0.0) {
    int q = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!(std::fabs(v[i]) <= 4.0)) return -1;
        for (;;) {
            const double s = std::ldexp(v[i], q);
            if (std::fabs(s - std::nearbyint(s)) <= std::ldexp(tol, q)) break;
            if (++q > qmax) return -1;
        }
    }
    return q;
}
// the integer m with v ~= m / 2^q, and whether the equality is exact
static long long dyadic_m(double v, int q, bool *exact) {
    const double s = std::ldexp(v, q), r = std::nearbyint(s);
    if (s != r) *exact = false;
    return (long long)r;
}
// libdevice's pow() may leave the bicubic weights a few ulps off their dyadic values; anything within 2^-44 keeps
// the FP64 sums of the reference within 1e-10 of the exact ones (see idot_round_u8 in vpp_march_kernels.cu)
constexpr double kDyadicTol = 1.0 / 17592186044416.0;  // 2^-44

// base-128 digits of the four weights of one dot product, packed one byte per tap (byte k <-> tap k)
static bool pack_digits(const long long m[4], bool lng, int w[4]) {
    unsigned u[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; ++k) {
        long long d[4];
        if (!lng) {
            d[0] = m[k] & 127;
            d[1] = m[k] >> 7;  // arithmetic
            d[2] = d[3] = 0;
            if (d[1] < -128 || d[1] > 127) return false;
        } else {
            const long long lo = m[k] & 16383, hi = m[k] >> 14;
            d[0] = lo & 127;
            d[1] = lo >> 7;
            d[2] = hi & 127;
            d[3] = hi >> 7;
            if (d[3] < -128 || d[3] > 127) return false;
        }
        for (int i = 0; i < 4; ++i) u[i] |= (unsigned)(d[i] & 0xff) << (8 * k);
    }
    for (int i = 0; i < 4; ++i) w[i] = (int)u[i];
    return true;
}

// n / D == (n * mul) >> (32 + shift) for every 0 <= n <= nmax ?  Checks every multiple of D and its predecessor.
static bool find_div_magic(unsigned D, unsigned long long nmax, unsigned *mul, int *shift) {
    if (D == 0) return false;
    if (D == 1) {  // n * (2^32 - 1) >> 32 = n - 1 for n > 0: handle as shift 0 with a 33-bit constant -> not representable
        return false;
    }
    int lg = 0;
    while ((2ull << lg) <= D) ++lg;  // floor(log2 D)
    for (int sh = 0; sh <= lg; ++sh) {
        const unsigned long long M = ((1ull << (32 + sh)) + D - 1) / D;  // ceil
        if (M >> 32) continue;
        bool good = true;
        for (unsigned long long n = 0; n <= nmax + D && good; n += D) {
            if (n > 0 && (((n - 1) * M) >> (32 + sh)) != (n - 1) / D) good = false;
            if (n <= nmax && ((n * M) >> (32 + sh)) != n / D) good = false;
        }
        if (good && ((nmax * M) >> (32 + sh)) == nmax / D) {
            *mul = (unsigned)M;
            *shift = sh;
            return true;
        }
    }
    return false;
}

static int build_march(const MarchInputs &in, int sm_count, MarchPlanHost &out) {
    (void)sm_count;
    const bool bicubic = in.rs == RS_BICUBIC, bilinear = in.rs == RS_BILINEAR;
    if (!bicubic && !bilinear && in.rs != RS_AREA) return TSVPP_UNSUPPORTED;
    const int BW = bilinear ? 4 : 2;  // output columns per consumer thread (march_bw)
    const int dw = in.dw, dh = in.dh, H = in.H;
    const Geometry &g = in.g;
    const bool need_uv = in.need_uv;

    // ---------------- exactness of the weights ----------------
    int qx = 0, qy = 0, ex = 0, ey = 0;
    bool xlong = false, ylong = false;
    std::vector<int> pxi, pyi;  // area integer patterns
    int NW = 1;
    std::vector<int> bmx, bmy;  // bilinear: weights of the second tap, m / 2^ex and m / 2^ey
    if (bilinear) {
        std::vector<double> vx(dw), vy(dh);
        for (int j = 0; j < dw; ++j) vx[j] = (*in.xs)[j].w;
        for (int i = 0; i < dh; ++i) vy[i] = (*in.ys)[i].w;
        ex = dyadic_q(vx.data(), vx.size(), 7);
        ey = dyadic_q(vy.data(), vy.size(), 7);
        if (ex < 0 || ey < 0) return TSVPP_UNSUPPORTED;  // 8 + ex + ey <= 22 bits: every f32 product and sum is exact
        bmx.resize(dw);
        bmy.resize(dh);
        for (int j = 0; j < dw; ++j) {
            bmx[j] = (int)std::ldexp(vx[j], ex);
            if (bmx[j] < 0 || bmx[j] > (1 << ex)) return TSVPP_UNSUPPORTED;
        }
        for (int i = 0; i < dh; ++i) {
            bmy[i] = (int)std::ldexp(vy[i], ey);
            if (bmy[i] < 0 || bmy[i] > (1 << ey)) return TSVPP_UNSUPPORTED;
        }
    } else if (bicubic) {
        qx = dyadic_q(in.cxw->data(), (size_t)4 * dw, 27, kDyadicTol);
        qy = dyadic_q(in.cyw->data(), (size_t)4 * dh, 27, kDyadicTol);
        if (qx < 0 || qy < 0) return TSVPP_UNSUPPORTED;
        if (qx < 1) qx = 1;
        if (qy < 1) qy = 1;
        if (qx > 13) { xlong = true; qx = std::max(qx, 15); }
        if (qy > 13) { ylong = true; qy = std::max(qy, 15); }
    } else {
        if (in.KX > 16 || in.K > 64) return TSVPP_UNSUPPORTED;
        std::vector<double> vx(in.patx->begin(), in.patx->end()), vy(in.paty->begin(), in.paty->end());
        ex = dyadic_q(vx.data(), vx.size(), 7);
        ey = dyadic_q(vy.data(), vy.size(), 7);
        if (ex < 0 || ey < 0) return TSVPP_UNSUPPORTED;
        pxi.resize(vx.size());
        pyi.resize(vy.size());
        for (size_t i = 0; i < vx.size(); ++i) {
            pxi[i] = (int)std::ldexp(vx[i], ex);
            if (pxi[i] < 0 || pxi[i] > 255) return TSVPP_UNSUPPORTED;
        }
        for (size_t i = 0; i < vy.size(); ++i) {
            pyi[i] = (int)std::ldexp(vy[i], ey);
            if (pyi[i] < 0 || pyi[i] > 255) return TSVPP_UNSUPPORTED;
        }
        NW = in.KX <= 4 ? 1 : in.KX <= 8 ? 2 : 4;
    }

    // ---------------- tiling ----------------
    // tile width tw (threads per group = tw/2), G row groups of rpg row pairs per strip
    const int K = in.K;
    auto luma_rows_of = [&](int i, int rows[64]) -> int {  // source rows of output luma row i, clamped like build_plan
        const AxisSample &s = (*in.ys)[i];
        int n = 0;
        if (bilinear) {
            int t[2];
            bilinear_taps(s.pos, 1, g.sh, t);
            for (int k = 0; k < 2; ++k) rows[n++] = std::max(0, std::min(g.y0 + t[k], H - 1));
        } else if (bicubic) {
            int t[4];
            bicubic_taps(s.pos, 1, g.sh, t);
            for (int k = 0; k < 4; ++k) rows[n++] = std::max(0, std::min(g.y0 + t[k], H - 1));
        } else {
            const int prow = i % in.PY;
            for (int k = 0; k < K; ++k)
                if (pyi[(size_t)prow * K + k] != 0) rows[n++] = std::max(0, std::min(g.y0 + s.pos + k, H - 1));
        }
        return n;
    };
    auto chroma_rows_of = [&](int ci, int rows[64]) -> int {
        const AxisSample &s = (*in.ys)[std::min(ci, dh - 1)];  // luma-grid formula at the chroma index
        int n = 0;
        if (bilinear) {
            int t[2];
            bilinear_taps(s.pos, 1, g.sh / 2, t);
            for (int k = 0; k < 2; ++k) rows[n++] = std::max(0, std::min(g.y0 / 2 + t[k], H / 2 - 1));
        } else if (bicubic) {
            int t[4];
            bicubic_taps(s.pos, 1, g.sh / 2, t);
            for (int k = 0; k < 4; ++k) rows[n++] = std::max(0, std::min(g.y0 / 2 + t[k], H / 2 - 1));
        } else {
            const int prow = ci % in.PY;
            for (int k = 0; k < K; ++k)
                if (pyi[(size_t)prow * K + k] != 0) rows[n++] = std::max(0, std::min(g.y0 / 2 + s.pos + k, H / 2 - 1));
        }
        return n;
    };
    // slot lists for a given strip height R (luma rows): returns max slots
    struct Slots {
        std::vector<std::vector<int>> l, c;
        int nsl = 0, nsc = 0;
    };
    auto make_slots = [&](int R, Slots &S) {
        const int nstrips = (dh + R - 1) / R;
        S.l.assign(nstrips, {});
        S.c.assign(nstrips, {});
        S.nsl = S.nsc = 1;
        int rows[64];
        for (int st = 0; st < nstrips; ++st) {
            std::vector<int> &L = S.l[st];
            for (int i = st * R; i < std::min(dh, (st + 1) * R); ++i) {
                const int n = luma_rows_of(i, rows);
                L.insert(L.end(), rows, rows + n);
            }
            std::sort(L.begin(), L.end());
            L.erase(std::unique(L.begin(), L.end()), L.end());
            S.nsl = std::max(S.nsl, (int)L.size());
            if (need_uv) {
                std::vector<int> &C = S.c[st];
                for (int ci = st * (R / 2); ci < std::min(in.ch2, (st + 1) * (R / 2)); ++ci) {
                    const int n = chroma_rows_of(ci, rows);
                    C.insert(C.end(), rows, rows + n);
                }
                std::sort(C.begin(), C.end());
                C.erase(std::unique(C.begin(), C.end()), C.end());
                S.nsc = std::max(S.nsc, (int)C.size());
            }
        }
    };

    // per-column window geometry relative to the tile segment start, for a candidate tile width
    const int lwords = bilinear ? 2 : bicubic ? 2 : NW + 1;      // 32-bit words a luma column reads per row
    const int cwords = bilinear ? 2 : bicubic ? 3 : 2 * NW + 1;  // words a chroma pair reads per row
    struct Tiles {
        std::vector<int> a0l, lenl, a0c, lenc;
        int stride = 16, nt = 1;
    };
    auto make_tiles = [&](int tw, Tiles &T) {
        T.nt = (dw + tw - 1) / tw;
        T.a0l.assign(T.nt, 0); T.lenl.assign(T.nt, 16); T.a0c.assign(T.nt, 0); T.lenc.assign(T.nt, 16);
        T.stride = 16;
        for (int t = 0; t < T.nt; ++t) {
            const int j0 = t * tw, j1 = std::min(dw, j0 + tw);
            int lo = 1 << 30, hi = -1;
            for (int j = j0; j < j1; ++j) {
                lo = std::min(lo, (*in.lo_l)[j]);
                hi = std::max(hi, (*in.hi_l)[j]);
            }
            T.a0l[t] = lo & ~15;
            int need = hi + 1 - T.a0l[t];
            T.lenl[t] = std::min(round_up(need, 16), round_up(in.pitch_y - T.a0l[t], 16));
            // the window reads whole words from the 4-byte-aligned start of the first tap
            int reach = 0;
            for (int j = j0; j < j1; ++j) reach = std::max(reach, (((*in.lo_l)[j] - T.a0l[t]) & ~3) + 4 * lwords);
            T.stride = std::max(T.stride, round_up(std::max(need, reach), 16));
            if (need_uv) {
                lo = 1 << 30;
                hi = -1;
                for (int b = j0; b < j1; ++b) {
                    if ((*in.lo_c)[b] < 0) continue;
                    lo = std::min(lo, (*in.lo_c)[b]);
                    hi = std::max(hi, (*in.hi_c)[b]);
                }
                if (hi < 0) { lo = 0; hi = 0; }
                T.a0c[t] = lo & ~15;
                need = hi + 1 - T.a0c[t];
                T.lenc[t] = std::min(round_up(need, 16), round_up(in.pitch_uv - T.a0c[t], 16));
                reach = 0;
                for (int b = j0; b < j1; b += 2) {
                    if ((*in.lo_c)[b] < 0) continue;
                    reach = std::max(reach, (((*in.lo_c)[b] - T.a0c[t]) & ~3) + 4 * cwords);
                }
                T.stride = std::max(T.stride, round_up(std::max(need, reach), 16));
            }
        }
    };

    int tw = 0, G = 1, rpg = 1, nstages = 2;
    Slots SL;
    Tiles TL;
    {
        long best = -1;
        const int want_tw = env_int("TSVPP_MARCH_TW", 0), want_g = env_int("TSVPP_MARCH_G", 0),
                  want_rpg = env_int("TSVPP_MARCH_RPG", 0), want_ns = env_int("TSVPP_MARCH_STAGES", 0);
        for (int ctw = 256; ctw >= 32; ctw >>= 1) {
            if (want_tw && ctw != want_tw) continue;
            if (ctw > 32 && ctw >= 2 * round_up(dw, 32)) continue;
            Tiles T;
            make_tiles(ctw, T);
            for (int crpg = 8; crpg >= 1; crpg >>= 1) {
                if (want_rpg && crpg != want_rpg) continue;
                for (int cG = 8; cG >= 1; cG >>= 1) {
                    if (want_g && cG != want_g) continue;
                    const int cons = (ctw / BW) * cG;
                    if (cons > kMarchMaxThreads - 32) continue;
                    const int R = 2 * cG * crpg;
                    if (R > 2 && R >= 2 * round_up(dh, 2)) continue;  // don't over-size strips for short outputs
                    Slots S;
                    make_slots(R, S);
                    const size_t stage = (size_t)(S.nsl + (need_uv ? S.nsc : 0)) * T.stride;
                    for (int ns = 3; ns >= 2; --ns) {
                        if (want_ns && ns != want_ns) continue;
                        const size_t smem = stage * ns + 2 * 8 * kMaxStages + 128;
                        if (smem > in.smem_cap) continue;
                        const int threads = 32 + round_up(cons, 32);
                        const int by_smem = (int)((size_t)(228 * 1024) / (smem + 1024));
                        const int by_regs = std::max(1, 65536 / (threads * 64));
                        const int ctas = std::max(1, std::min(std::min(by_smem, by_regs), 2048 / threads));
                        // useful fraction of the marched rows: a group restarts its window (3 extra source rows)
                        double eff = 1.0;
                        if (bilinear) {
                            const double rows_ideal = 2.0 * crpg * (double)g.sh / dh;
                            eff = rows_ideal / (rows_ideal + 1.0);
                        }
                        if (bicubic) {
                            const double rows_ideal = 2.0 * crpg * std::max(1.0, (double)g.sh / dh);
                            eff = rows_ideal / (rows_ideal + 3.0);
                        }
                        const long resident = (long)ctas * cons;
                        const long score = (long)(std::min(resident, 1536L) * eff * 64) + (ctw / 32) * 8 + (ns == 2 ? 2 : 0) +
                                           (cons >= 64 ? 64 : 0);
                        if (score > best) {
                            best = score;
                            tw = ctw; G = cG; rpg = crpg; nstages = ns;
                            SL = S;
                            TL = T;
                        }
                    }
                }
            }
        }
        if (best < 0) return TSVPP_UNSUPPORTED;
    }
    const int R = 2 * G * rpg;
    const int nstrips = (dh + R - 1) / R;
    const int ntiles = TL.nt;
    const int nsl = SL.nsl, nsc = need_uv ? SL.nsc : 0;
    const int slot_stride = TL.stride;

    // ---------------- row tables ----------------
    auto slot_of = [](const std::vector<int> &list, int row) -> int {
        return (int)(std::lower_bound(list.begin(), list.end(), row) - list.begin());
    };
    DeviceArena ar;
    size_t off_row_l = 0, off_row_c = 0, off_gfl = 0, off_gfc = 0;
    int rows[64];
    if (bilinear) {
        std::vector<MarchRowL> rl((size_t)nstrips * R + 2, MarchRowL{0}), rc((size_t)std::max(1, nstrips * (R / 2)) + 2, MarchRowL{0});
        std::vector<int> gfl((size_t)nstrips * G, 0), gfc((size_t)nstrips * G, 0);
        // the two-row window {marched - 1, marched}; taps are folded onto it like the bicubic ones (the bottom
        // border pulls a row's taps back: yDiff = 0, src/Resize.cu:9-10)
        auto fill_row = [&](MarchRowL &r, const std::vector<int> &list, const int *srows, int my, int &marched, int *first_out) -> bool {
            const int s0 = slot_of(list, srows[0]), s1 = slot_of(list, srows[1]);
            marched = std::max(marched, std::max(s0, s1));
            if (marched - std::min(s0, s1) > 1) return false;
            int w[2] = {0, 0};  // [0] = slot marched - 1, [1] = slot marched
            w[1 - (marched - s0)] += (1 << ey) - my;
            w[1 - (marched - s1)] += my;
            if (first_out) *first_out = std::min(s0, s1);
            if (marched > 0xffff || w[0] > 255 || w[1] > 255) return false;
            r.packed = (unsigned)marched | ((unsigned)w[0] << 16) | ((unsigned)w[1] << 24);
            return true;
        };
        for (int st = 0; st < nstrips; ++st) {
            int last_valid = -1, marched = -1;
            for (int k = 0; k < R; ++k) {
                const int i = st * R + k;
                const bool gstart = k % (2 * rpg) == 0;
                if (gstart) marched = -1;
                if (i < dh) {
                    luma_rows_of(i, rows);
                    if (!fill_row(rl[i], SL.l[st], rows, bmy[i], marched, gstart ? &gfl[(size_t)st * G + k / (2 * rpg)] : nullptr))
                        return TSVPP_UNSUPPORTED;
                    last_valid = i;
                } else {
                    rl[i] = rl[last_valid];
                }
            }
            if (need_uv) {
                last_valid = -1;
                marched = -1;
                for (int k = 0; k < R / 2; ++k) {
                    const int ci = st * (R / 2) + k;
                    const bool gstart = k % rpg == 0;
                    if (gstart) marched = -1;
                    if (ci < in.ch2) {
                        chroma_rows_of(ci, rows);
                        // the chroma row is weighted like luma row index ci (same x, y, wx, wy: src/Resize.cu:303-310)
                        if (!fill_row(rc[ci], SL.c[st], rows, bmy[std::min(ci, dh - 1)], marched, gstart ? &gfc[(size_t)st * G + k / rpg] : nullptr))
                            return TSVPP_UNSUPPORTED;
                        last_valid = ci;
                    } else if (last_valid >= 0) {
                        rc[ci] = rc[last_valid];
                    }
                }
            }
        }
        off_row_l = ar.add(rl.data(), rl.size() * sizeof(MarchRowL));
        off_row_c = ar.add(rc.data(), rc.size() * sizeof(MarchRowL));
        off_gfl = ar.add(gfl.data(), gfl.size() * sizeof(int));
        off_gfc = ar.add(gfc.data(), gfc.size() * sizeof(int));
    } else if (bicubic) {
        std::vector<MarchRowB> rl((size_t)nstrips * R), rc((size_t)std::max(1, nstrips * (R / 2)));
        std::memset(rl.data(), 0, rl.size() * sizeof(MarchRowB));
        std::memset(rc.data(), 0, rc.size() * sizeof(MarchRowB));
        std::vector<int> gfl((size_t)nstrips * G, 0), gfc((size_t)nstrips * G, 0);
        // `marched`: how far the window of the row's group has already been marched (-1 at the start of a group).
        // At the bottom border the reference pulls its taps back (yDiff = 0 when y+2 >= height, src/Resize.cu:36-39),
        // so a row's newest tap can lie BEHIND the previous row's: the window stays where it is and the taps are
        // folded onto its bytes relative to the marched position (byte 0 = the marched slot).
        auto fill_row = [&](MarchRowB &r, const std::vector<int> &list, const int *srows, const double *cw, int q,
                            bool lng, int &marched, int *first_out) -> bool {
            int s[4];
            for (int k = 0; k < 4; ++k) s[k] = slot_of(list, srows[k]);
            const int oldest = std::min(std::min(s[0], s[1]), std::min(s[2], s[3]));
            marched = std::max(marched, std::max(std::max(s[0], s[1]), std::max(s[2], s[3])));
            if (marched - oldest > 3 || marched > 0xfff) return false;
            long long m[4] = {0, 0, 0, 0};
            bool exact = true;
            unsigned tapsel = 0;
            for (int k = 0; k < 4; ++k) {
                const int byte = marched - s[k];
                m[byte] += dyadic_m(cw[k], q, &exact);
                tapsel |= (unsigned)byte << (4 * k);
            }
            if (first_out) *first_out = oldest;
            r.meta = (unsigned)marched | (tapsel << 12) | (exact ? 1u << 28 : 0u);
            return pack_digits(m, lng, r.w);
        };
        for (int st = 0; st < nstrips; ++st) {
            int last_valid = -1, marched = -1;
            for (int k = 0; k < R; ++k) {
                const int i = st * R + k;
                const bool gstart = k % (2 * rpg) == 0;
                if (gstart) marched = -1;  // a new row group starts its own window
                if (i < dh) {
                    luma_rows_of(i, rows);
                    if (!fill_row(rl[i], SL.l[st], rows, in.cyw->data() + (size_t)4 * i, qy, ylong, marched,
                                  gstart ? &gfl[(size_t)st * G + k / (2 * rpg)] : nullptr))
                        return TSVPP_UNSUPPORTED;
                    last_valid = i;
                } else {
                    rl[i] = rl[last_valid];  // never emitted; keeps the march where it is
                }
            }
            if (need_uv) {
                last_valid = -1;
                marched = -1;
                for (int k = 0; k < R / 2; ++k) {
                    const int ci = st * (R / 2) + k;
                    const bool gstart = k % rpg == 0;
                    if (gstart) marched = -1;
                    if (ci < in.ch2) {
                        chroma_rows_of(ci, rows);
                        // vertical weights of chroma row ci are those of luma row index ci (src/Resize.cu:351-353)
                        if (!fill_row(rc[ci], SL.c[st], rows, in.cyw->data() + (size_t)4 * std::min(ci, dh - 1), qy, ylong, marched,
                                      gstart ? &gfc[(size_t)st * G + k / rpg] : nullptr))
                            return TSVPP_UNSUPPORTED;
                        last_valid = ci;
                    } else if (last_valid >= 0) {
                        rc[ci] = rc[last_valid];
                    }
                }
            }
        }
        off_row_l = ar.add(rl.data(), rl.size() * sizeof(MarchRowB));
        off_row_c = ar.add(rc.data(), rc.size() * sizeof(MarchRowB));
        off_gfl = ar.add(gfl.data(), gfl.size() * sizeof(int));
        off_gfc = ar.add(gfc.data(), gfc.size() * sizeof(int));
    } else {
        std::vector<MarchRowA> rl((size_t)nstrips * R, MarchRowA{0, 0}), rc((size_t)std::max(1, nstrips * (R / 2)), MarchRowA{0, 0});
        for (int st = 0; st < nstrips; ++st) {
            for (int k = 0; k < R; ++k) {
                const int i = st * R + k;
                if (i >= dh) continue;
                const int n = luma_rows_of(i, rows);
                if (n <= 0) return TSVPP_UNSUPPORTED;
                const int f = slot_of(SL.l[st], rows[0]);
                for (int a = 0; a < n; ++a)  // the non-zero-weight rows must sit in consecutive slots
                    if (slot_of(SL.l[st], rows[a]) != f + a) return TSVPP_UNSUPPORTED;
                // ... and be the FIRST n window rows (zero padding only at the end of a pattern row)
                for (int a = 0; a < n; ++a)
                    if (pyi[(size_t)(i % in.PY) * K + a] == 0) return TSVPP_UNSUPPORTED;
                rl[i] = MarchRowA{f, i % in.PY};
            }
            if (need_uv)
                for (int k = 0; k < R / 2; ++k) {
                    const int ci = st * (R / 2) + k;
                    if (ci >= in.ch2) continue;
                    const int n = chroma_rows_of(ci, rows);
                    if (n <= 0) return TSVPP_UNSUPPORTED;
                    const int f = slot_of(SL.c[st], rows[0]);
                    for (int a = 0; a < n; ++a)
                        if (slot_of(SL.c[st], rows[a]) != f + a) return TSVPP_UNSUPPORTED;
                    rc[ci] = MarchRowA{f, ci % in.PY};
                }
        }
        off_row_l = ar.add(rl.data(), rl.size() * sizeof(MarchRowA));
        off_row_c = ar.add(rc.data(), rc.size() * sizeof(MarchRowA));
    }

    // ---------------- column tables ----------------
    const int ncol = ntiles * tw;
    size_t off_col_l = 0, off_col_c = 0, off_pxw = 0, off_pyw = 0, off_dmul = 0, off_dsh = 0;
    if (bilinear) {
        std::vector<MarchColL> cl(ncol, MarchColL{0, 0x3210u, 0x3210u, 0}), cc(ncol / 2 + 2, MarchColL{0, 0x3210u, 0x3210u, 0});
        auto weights = [&](int mx, bool same) -> unsigned {
            // taps that coincide (right border, xDiff = 0) carry the whole weight on byte 0
            return same ? (unsigned)(1 << ex) : ((unsigned)((1 << ex) - mx) | ((unsigned)mx << 8));
        };
        for (int j = 0; j < dw; ++j) {
            const int t = j / tw;
            const int o0 = (*in.tap_l)[(size_t)j * 4] - TL.a0l[t], o1 = (*in.tap_l)[(size_t)j * 4 + 1] - TL.a0l[t];
            if (o0 < 0 || o1 < o0) return TSVPP_UNSUPPORTED;
            const int base = o0 & ~3;
            if (o1 - base > 7) return TSVPP_UNSUPPORTED;
            cl[j].addr = base;
            cl[j].sel0 = (unsigned)(o0 - base) | ((unsigned)(o1 - base) << 4) | 0x3200u;
            cl[j].w = weights(bmx[j], o1 == o0);
        }
        if (need_uv)
            for (int cj = 0; cj < in.cw2; ++cj) {
                const int j = 2 * cj, t = j / tw;
                const int u0 = (*in.tap_c)[(size_t)j * 4] - TL.a0c[t], u1 = (*in.tap_c)[(size_t)j * 4 + 1] - TL.a0c[t];
                const int v0 = (*in.tap_c)[(size_t)(j + 1) * 4] - TL.a0c[t], v1 = (*in.tap_c)[(size_t)(j + 1) * 4 + 1] - TL.a0c[t];
                if (u0 < 0 || u1 < u0 || v0 < u0 || v1 < v0) return TSVPP_UNSUPPORTED;
                const int base = u0 & ~3;
                if (v1 - base > 7 || u1 - base > 7) return TSVPP_UNSUPPORTED;
                if ((u1 == u0) != (v1 == v0)) return TSVPP_UNSUPPORTED;
                cc[cj].addr = base;
                cc[cj].sel0 = (unsigned)(u0 - base) | ((unsigned)(u1 - base) << 4) | 0x3200u;
                cc[cj].sel1 = (unsigned)(v0 - base) | ((unsigned)(v1 - base) << 4) | 0x3200u;
                cc[cj].w = weights(bmx[cj], u1 == u0);  // the pair is weighted like luma column cj
            }
        off_col_l = ar.add(cl.data(), cl.size() * sizeof(MarchColL));
        off_col_c = ar.add(cc.data(), cc.size() * sizeof(MarchColL));
    } else if (bicubic) {
        MarchColB z;
        std::memset(&z, 0, sizeof(z));
        z.sel0 = 0x3210u;
        std::vector<MarchColB> cl(ncol, z), cc(ncol / 2 + 1, z);
        for (int j = 0; j < dw; ++j) {
            const int t = j / tw;
            int o[4], mn = 1 << 30;
            for (int k = 0; k < 4; ++k) {
                o[k] = (*in.tap_l)[(size_t)j * 4 + k] - TL.a0l[t];
                if (o[k] < 0) return TSVPP_UNSUPPORTED;
                mn = std::min(mn, o[k]);
            }
            const int base = mn & ~3;
            unsigned sel = 0;
            long long m[4];
            bool exact = true;
            for (int k = 0; k < 4; ++k) {
                const int b = o[k] - base;
                if (b > 7) return TSVPP_UNSUPPORTED;
                sel |= (unsigned)b << (4 * k);
                m[k] = dyadic_m((*in.cxw)[(size_t)4 * j + k], qx, &exact);
            }
            cl[j].addr = base;
            cl[j].sel0 = sel;
            cl[j].exact = exact ? 1 : 0;
            if (!pack_digits(m, xlong, cl[j].w)) return TSVPP_UNSUPPORTED;
        }
        if (need_uv)
            for (int cj = 0; cj < in.cw2; ++cj) {
                const int j = 2 * cj, t = j / tw;
                int mn = 1 << 30, o[2][4];
                for (int comp = 0; comp < 2; ++comp)
                    for (int k = 0; k < 4; ++k) {
                        o[comp][k] = (*in.tap_c)[(size_t)(j + comp) * 4 + k] - TL.a0c[t];
                        if (o[comp][k] < 0) return TSVPP_UNSUPPORTED;
                        mn = std::min(mn, o[comp][k]);
                    }
                const int base = mn & ~3;
                unsigned selA[2] = {0, 0}, selB[2] = {0, 0};
                for (int comp = 0; comp < 2; ++comp)
                    for (int k = 0; k < 4; ++k) {
                        const int b = o[comp][k] - base;
                        if (b > 11) return TSVPP_UNSUPPORTED;
                        if (b < 8) {
                            selA[comp] |= (unsigned)b << (4 * k);
                            selB[comp] |= (unsigned)k << (4 * k);  // keep byte k of the first pick
                        } else {
                            selB[comp] |= (unsigned)(4 + (b - 8)) << (4 * k);  // byte of the third word
                        }
                    }
                long long m[4];
                bool exact = true;
                for (int k = 0; k < 4; ++k) m[k] = dyadic_m((*in.cxw)[(size_t)4 * cj + k], qx, &exact);  // weights of xs[cj]
                cc[cj].addr = base;
                cc[cj].exact = exact ? 1 : 0;
                cc[cj].sel0 = selA[0] | (selA[1] << 16);
                cc[cj].sel1 = selB[0] | (selB[1] << 16);
                if (!pack_digits(m, xlong, cc[cj].w)) return TSVPP_UNSUPPORTED;
            }
        off_col_l = ar.add(cl.data(), cl.size() * sizeof(MarchColB));
        off_col_c = ar.add(cc.data(), cc.size() * sizeof(MarchColB));
    } else {
        std::vector<MarchColA> cl(ncol, MarchColA{0, 0, 0, 0}), cc(ncol / 2 + 1, MarchColA{0, 0, 0, 0});
        for (int j = 0; j < dw; ++j) {
            const int t = j / tw;
            const int o = (*in.tap_l)[(size_t)j * 4] - TL.a0l[t];
            if (o < 0) return TSVPP_UNSUPPORTED;
            cl[j] = MarchColA{o & ~3, 8 * (o & 3), j % in.PX, 0};
        }
        if (need_uv)
            for (int cj = 0; cj < in.cw2; ++cj) {
                const int j = 2 * cj, t = j / tw;
                const int o = (*in.tap_c)[(size_t)j * 4] - TL.a0c[t];
                if (o < 0 || (*in.tap_c)[(size_t)(j + 1) * 4] != (*in.tap_c)[(size_t)j * 4] + 1) return TSVPP_UNSUPPORTED;
                cc[cj] = MarchColA{o & ~3, 8 * (o & 3), cj % in.PX, 0};
            }
        off_col_l = ar.add(cl.data(), cl.size() * sizeof(MarchColA));
        off_col_c = ar.add(cc.data(), cc.size() * sizeof(MarchColA));
        // packed horizontal weights, integer vertical weights, divisors
        std::vector<int> pxw((size_t)in.PX * NW, 0);
        for (int r = 0; r < in.PX; ++r)
            for (int b = 0; b < in.KX; ++b)
                pxw[(size_t)r * NW + b / 4] |= (int)((unsigned)pxi[(size_t)r * in.KX + b] << (8 * (b % 4)));
        std::vector<unsigned> dmul((size_t)in.PX * in.PY);
        std::vector<int> dsh((size_t)in.PX * in.PY);
        for (int r = 0; r < in.PY; ++r)
            for (int c = 0; c < in.PX; ++c) {
                unsigned long long D = 0;
                float divf = 0.f;  // the reference's own accumulation (src/Resize.cu:166-173)
                for (int a = 0; a < K; ++a)
                    for (int b = 0; b < in.KX; ++b) {
                        D += (unsigned long long)pxi[(size_t)c * in.KX + b] * pyi[(size_t)r * K + a];
                        divf = divf + (*in.patx)[(size_t)c * in.KX + b] * (*in.paty)[(size_t)r * K + a];
                    }
                // exact f32 accumulation of every partial sum, and a quotient that cannot round up to the next integer
                if (D == 0 || 255ull * D >= (1ull << 23)) return TSVPP_UNSUPPORTED;
                if ((double)divf != std::ldexp((double)D, -(ex + ey))) return TSVPP_UNSUPPORTED;
                if (!find_div_magic((unsigned)D, 255ull * D, &dmul[(size_t)r * in.PX + c], &dsh[(size_t)r * in.PX + c]))
                    return TSVPP_UNSUPPORTED;
            }
        off_pxw = ar.add(pxw.data(), pxw.size() * sizeof(int));
        off_pyw = ar.add(pyi.data(), pyi.size() * sizeof(int));
        off_dmul = ar.add(dmul.data(), dmul.size() * sizeof(unsigned));
        off_dsh = ar.add(dsh.data(), dsh.size() * sizeof(int));
    }

    size_t off_cxd = 0, off_cyd = 0;
    if (bicubic) {
        std::vector<double> cxd((size_t)4 * (dw + 2), 0.0);
        std::memcpy(cxd.data(), in.cxw->data(), (size_t)4 * dw * sizeof(double));
        off_cxd = ar.add(cxd.data(), cxd.size() * sizeof(double));
        std::vector<double> cyd((size_t)4 * (dh + 2), 0.0);  // rows dh, dh+1: the never-emitted second row of an odd last pair
        std::memcpy(cyd.data(), in.cyw->data(), (size_t)4 * dh * sizeof(double));
        off_cyd = ar.add(cyd.data(), cyd.size() * sizeof(double));
    }

    // ---------------- slot rows, segments ----------------
    std::vector<int> srl((size_t)nstrips * nsl, 0), src_((size_t)nstrips * std::max(1, nsc), 0), nl(nstrips, 0), nc(nstrips, 0);
    for (int st = 0; st < nstrips; ++st) {
        nl[st] = (int)SL.l[st].size();
        for (int k = 0; k < nl[st]; ++k) srl[(size_t)st * nsl + k] = SL.l[st][k];
        if (need_uv) {
            nc[st] = (int)SL.c[st].size();
            for (int k = 0; k < nc[st]; ++k) src_[(size_t)st * nsc + k] = SL.c[st][k];
        }
    }
    const size_t off_srl = ar.add(srl.data(), srl.size() * sizeof(int));
    const size_t off_src = ar.add(src_.data(), src_.size() * sizeof(int));
    const size_t off_nl = ar.add(nl.data(), nl.size() * sizeof(int));
    const size_t off_nc = ar.add(nc.data(), nc.size() * sizeof(int));
    const size_t off_a0l = ar.add(TL.a0l.data(), TL.a0l.size() * sizeof(int));
    const size_t off_lenl = ar.add(TL.lenl.data(), TL.lenl.size() * sizeof(int));
    const size_t off_a0c = ar.add(TL.a0c.data(), TL.a0c.size() * sizeof(int));
    const size_t off_lenc = ar.add(TL.lenc.data(), TL.lenc.size() * sizeof(int));

    TSVPP_CUDA(cudaMalloc(&out.dev_tables, ar.host.size() + 256));
    TSVPP_CUDA(cudaMemcpy(out.dev_tables, ar.host.data(), ar.host.size(), cudaMemcpyHostToDevice));
    const unsigned char *base = (const unsigned char *)out.dev_tables;

    MarchPlanDev &D = out.dev;
    std::memset(&D, 0, sizeof(D));
    D.e.dw = dw; D.e.dh = dh; D.e.cw2 = in.cw2; D.e.ch2 = in.ch2; D.e.swap_rb = in.swap_rb; D.e.need_uv = need_uv ? 1 : 0;
    D.mode = bilinear ? MARCH_BILINEAR : bicubic ? MARCH_BICUBIC : MARCH_AREA;
    D.vshift = ex + ey;
    D.tw = tw; D.ntiles = ntiles; D.nstrips = nstrips; D.G = G; D.rpg = rpg;
    D.nsl = nsl; D.nsc = nsc; D.slot_stride = slot_stride;
    D.stage_bytes = (nsl + nsc) * slot_stride;
    D.nstages = nstages;
    D.pitch_y = in.pitch_y; D.pitch_uv = in.pitch_uv; D.need_uv = need_uv ? 1 : 0;
    D.xlong = xlong ? 1 : 0; D.ylong = ylong ? 1 : 0; D.qx = qx; D.qy = qy;
    D.K = K; D.KX = in.KX; D.NW = NW; D.PX = in.PX; D.PY = in.PY;
    D.seg_a0_l = (const int *)(base + off_a0l); D.seg_len_l = (const int *)(base + off_lenl);
    D.seg_a0_c = (const int *)(base + off_a0c); D.seg_len_c = (const int *)(base + off_lenc);
    D.slotrow_l = (const int *)(base + off_srl); D.slotrow_c = (const int *)(base + off_src);
    D.nslots_l = (const int *)(base + off_nl); D.nslots_c = (const int *)(base + off_nc);
    D.col_l = base + off_col_l; D.col_c = base + off_col_c;
    D.row_l = base + off_row_l; D.row_c = base + off_row_c;
    D.gfirst_l = (const int *)(base + off_gfl); D.gfirst_c = (const int *)(base + off_gfc);
    D.cxd = (const double *)(base + off_cxd); D.cyd = (const double *)(base + off_cyd);
    D.px_w = (const int *)(base + off_pxw); D.py_w = (const int *)(base + off_pyw);
    D.div_mul = (const unsigned *)(base + off_dmul); D.div_shift = (const int *)(base + off_dsh);

    out.smem = march_smem_bytes(D);
    if (out.smem > in.smem_cap) return TSVPP_UNSUPPORTED;
    out.threads = march_block_threads(D);
    int rc = configure_march(D, in.ok, in.smem_cap);
    if (rc) return rc;
    out.ctas_per_sm = std::max(1, march_max_ctas_per_sm(D, in.ok, out.threads, out.smem));
    long fetch = 0;
    for (int st = 0; st < nstrips; ++st)
        for (int t = 0; t < ntiles; ++t) fetch += (long)nl[st] * TL.lenl[t] + (need_uv ? (long)nc[st] * TL.lenc[t] : 0);
    out.fetch_bytes_per_frame = fetch;
    char buf[640];
    std::snprintf(buf, sizeof(buf),
                  "kernel=march resample=%s-int out_kind=%d fill=bulk bw=%d tw=%d G=%d rpg=%d strip_rows=%d ntiles=%d nstrips=%d "
                  "slots=%d+%d slot_stride=%d stage_bytes=%d nstages=%d smem=%zu threads=%d ctas_per_sm=%d qx=%d qy=%d "
                  "digits=%s/%s K=%d KX=%d NW=%d PX=%d PY=%d crop=%d resize=1 fetch_bytes_per_frame=%ld direct=0",
                  bilinear ? "bilinear" : bicubic ? "bicubic" : "area", in.ok, BW, tw, G, rpg, R, ntiles, nstrips, nsl, nsc, slot_stride, D.stage_bytes,
                  nstages, out.smem, out.threads, out.ctas_per_sm, bicubic ? qx : ex, bicubic ? qy : ey,
                  xlong ? "long" : "short", ylong ? "long" : "short", K, in.KX, NW, in.PX, in.PY, (int)g.crop, fetch);
    out.desc = buf;
    return TSVPP_OK;
}
