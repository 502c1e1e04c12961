/* TiktakGlobalSearch -- host driver with the reference's process surface, on top of libtiktak_cuda.
 *
 *     ./TiktakGlobalSearch <-1|0|1|2|3|4|5> configfile <a|b|d>        (README.md:50-70)
 *
 * Mirrors, for one instance on one GPU: parseCommandLine (TiktakGlobalSearch.f90:1150-1246), parseConfig
 * (genericParams.f90:265-450), the state sequence 1->2->3->(4)->6->7->(8)->10 of a cold start (:217-401)
 * and the output files with their record formats:
 *   sobol.dat / x_starts.dat   mywrite2  '(1x,1000(3x,F30.15))'   (stateControl.f90:283)
 *   sobolFnVal.dat             '(i8, 200f40.20)'                   (TiktakGlobalSearch.f90:467)
 *   searchStart.dat            '(2i10, 200f40.20)'                 (:598)
 *   searchResults.dat          '(3i10, 200f40.20)'                 (:597)
 *   FinalStart.dat / FinalResults.dat '(2i10, 200f40.20)'          (:677)
 *   state.dat, seqNo.dat, lastSobol.dat, legitSobol.dat, lastParam.dat  list-directed integers
 * The reference's Fortran toolchain is not in the build image, so this driver is C++; the Fortran
 * binding of the same ABI is in INTEGRATION.md.  Options 1/-1 coordinate several CPU processes through lock files in
 * the reference; with device-resident state there is nothing to join, so they only report that.
 * Options 2 and 3 (genericParams.f90:118-251) resume from the files of an earlier run:
 *   2  more Sobol points / a higher `legitimate`: the same consistency checks against internalConfig.dat,
 *      sobolFnVal.dat and legitSobol.dat, then the Sobol stage for the new config (the rows already in sobolFnVal.dat
 *      are re-derived bit for bit -- evaluating them again costs microseconds, parsing them would not be faster) and
 *      everything downstream from scratch, as the reference does (it resets searchResults.dat etc., :118-131).
 *   3  more local searches: same Sobol config required; the finished rows of searchResults.dat are loaded back into
 *      the device-resident results table (tiktak_cuda_push_results) and the restarts continue at lastParam + 1,
 *      appending to searchStart.dat / searchResults.dat.
 * Knobs outside the reference surface come from the environment: TIKTAK_OBJECTIVE (0 template, 1 Griewank,
 * 2 Rastrigin, 3 Rosenbrock), TIKTAK_ROUND_SIZE (default 1 = the sequential reference order), TIKTAK_DEVICE,
 * TIKTAK_WRITE_SOBOL (1 = also dump sobol.dat like setupSobol does), TIKTAK_NGPU (default 1: shard the Sobol index range and
 * the restarts of every launch over the first n GPUs of the box; the library's NCCL communicator does the exchanges,
 * one host thread per GPU makes the stage calls, this thread -- rank 0 -- also writes the files).
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../include/tiktak_cuda.h"

static void usage() { /* myusage, :1235-1246 */
    printf("Usage: TiktakGlobalSearch <-1|0|1|2|3|4|5> [configfile] [a|b|d]\n"
           "  -1 exit state, 0 cold start, 1 warm start, 2 update sobol points, 3 update local searches,\n"
           "   4 diagnostic (objective at the initial guess), 5 local minimization from the initial guess\n"
           "  a = amoeba (Nelder-Mead), b = DFNLS (default), d = dfpmin (BFGS)\n");
}
static bool next_data_line(std::ifstream& f, std::string& line) {
    while (std::getline(f, line)) { if (!line.empty() && line[0] == '!') continue; return true; }
    return false;
}
static std::vector<double> nums(const std::string& line, size_t n) {
    std::string s = line.substr(0, line.find('!'));
    for (char& c : s) { if (c == ',') c = ' '; if (c == 'D' || c == 'd') c = 'e'; }
    std::istringstream is(s);
    std::vector<double> v; double x;
    while (v.size() < n && (is >> x)) v.push_back(x);
    return v;
}
struct Cfg { int nx, nm, maxeval, qr, legit, maxpoints, stype, lottery; double fvalmax; std::vector<double> range, init, bound; };
static bool parse_config(const char* path, Cfg& c) { /* parseConfig, genericParams.f90:265-450 */
    std::ifstream f(path);
    if (!f) { printf("<genericParams.parseConfig()> Error: unable to open file:%s\n", path); return false; }
    std::string line;
    double sc[9];
    if (!next_data_line(f, line) || nums(line, 1).empty()) goto bad;
    sc[0] = nums(line, 1)[0];
    for (int i = 1; i < 9; i++) { if (!std::getline(f, line) || nums(line, 1).empty()) goto bad; sc[i] = nums(line, 1)[0]; }
    c.nx = (int)sc[0]; c.nm = (int)sc[1]; c.maxeval = (int)sc[2]; c.qr = (int)sc[3]; c.legit = (int)sc[4]; c.fvalmax = sc[5];
    c.maxpoints = (int)sc[6]; c.stype = (int)sc[7]; c.lottery = (int)sc[8];
    if (tiktak_cuda_config_defaults(c.nx, &c.maxeval, &c.qr, &c.legit, &c.fvalmax, &c.maxpoints, &c.lottery)) {
        printf("<genericParams:parseConfig> Error in config file. maxpoints > # sobol points\n"); return false;
    }
    c.range.assign(2 * c.nx, 0); c.init.assign(c.nx, 0); c.bound.assign(2 * c.nx, 0);
    for (int grp = 0; grp < 3; grp++) {
        for (int i = 0; i < c.nx; i++) {
            if (!(i == 0 ? next_data_line(f, line) : (bool)std::getline(f, line))) goto bad;
            std::vector<double> v = nums(line, grp == 1 ? 1 : 2);
            if (v.size() < (size_t)(grp == 1 ? 1 : 2)) goto bad;
            if (grp == 0) { c.range[i] = v[0]; c.range[c.nx + i] = v[1]; }      /* column-major (nx,2) */
            else if (grp == 1) c.init[i] = v[0];
            else { c.bound[i] = v[0]; c.bound[c.nx + i] = v[1]; }
        }
    }
    return true;
bad:
    printf("<genericParams:parseConfig> Error. %s Unable to read all parameters\n", path);
    return false;
}
static void write_int(const char* file, long v) { FILE* f = fopen(file, "w"); if (f) { fprintf(f, " %11ld\n", v); fclose(f); } }
static void logmsg(const std::string& m) { FILE* f = fopen("logFile.txt", "a"); if (f) { fprintf(f, " %s\n", m.c_str()); fclose(f); } printf(" %s\n", m.c_str()); }
static void mywrite2(const char* file, const double* a, int rows, int cols) { /* a is column-major (rows, cols) */
    FILE* f = fopen(file, "w");
    if (!f) return;
    for (int r = 0; r < rows; r++) { fputc(' ', f); for (int c = 0; c < cols; c++) fprintf(f, "   %30.15f", a[(size_t)c * rows + r]); fputc('\n', f); }
    fclose(f);
}
static void broadcast_exit() { /* exitState, stateControl.f90:203-216: state -1 and poisoned work counters stop every instance on the directory */
    write_int("state.dat", -1);
    write_int("lastSobol.dat", -10000000);
    write_int("lastParam.dat", -10000000);
    write_int("seqNo.dat", -10000000);
}
static int exit_state(const std::string& m) {
    logmsg("EXIT STATE: " + m);
    broadcast_exit();
    return 1;
}
struct Cfg;
static void write_internal_config(const Cfg& c);
#define CHECK(call) do { int rc_ = (call); if (rc_) { char b_[1200]; snprintf(b_, sizeof b_, "EXIT STATE: %s: %s; %s", #call, tiktak_cuda_strerror(rc_), tiktak_cuda_last_error()); \
    logmsg(b_); broadcast_exit(); return 1; } } while (0)

static void write_internal_config(const Cfg& c) { /* genericParams.f90:154-174; parse_config reads it back (maxpoints is stored raw) */
    FILE* f = fopen("internalConfig.dat", "w");
    if (!f) return;
    fprintf(f, " %11d\n %11d\n %11d\n %11d\n %11d\n %24.16E\n %11d\n %11d\n %11d\n", c.nx, c.nm, c.maxeval, c.qr, c.legit, c.fvalmax, c.maxpoints, c.stype, c.lottery);
    for (int i = 0; i < c.nx; i++) fprintf(f, " %24.16E %24.16E\n", c.range[i], c.range[c.nx + i]);
    for (int i = 0; i < c.nx; i++) fprintf(f, " %24.16E\n", c.init[i]);
    for (int i = 0; i < c.nx; i++) fprintf(f, " %24.16E %24.16E\n", c.bound[i], c.bound[c.nx + i]);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 2) { usage(); return 0; }
    const int option = atoi(argv[1]);
    int alg = 0; /* p_default_alg = 0 = DFNLS (genericParams.f90:12) */
    const char* config = argc >= 3 ? argv[2] : "config.txt";
    const char* algarg = option == 1 ? (argc >= 3 ? argv[2] : nullptr) : (argc >= 4 ? argv[3] : nullptr); /* warm start: alg is arg 2 */
    if (algarg) { if (algarg[0] == 'a') alg = 1; else if (algarg[0] == 'b') alg = 0; else if (algarg[0] == 'd') alg = 2; else { usage(); return 0; } }
    if (option < -1 || option > 5) { usage(); return 0; }
    if (option == -1) { /* :219-223: tell every instance that shares the directory (CPU instances of the reference included) to stop */
        broadcast_exit();
        logmsg("exit state written: state.dat = -1, lastSobol/lastParam/seqNo poisoned");
        return 0;
    }
    if (option == 1) {
        logmsg("option 1 joins a multi-process CPU run through the .dat lock files; the GPU library keeps that state "
               "on the device and runs the whole job in one process (options 0, 2, 3). Nothing to do.");
        return 0;
    }
    Cfg c;
    if (!parse_config(config, c)) return 1;
    /* ---- options 2 / 3: consistency checks of initialize (genericParams.f90:118-251) ---- */
    std::vector<std::vector<double>> old_rows; /* searchResults.dat rows with k >= 1 (option 3) */
    int resume_k = 0;
    if (option == 2 || option == 3) {
        Cfg oldc;
        if (!parse_config("internalConfig.dat", oldc)) return exit_state("update: internalConfig.dat of the earlier run cannot be read");
        long numsolved = -1, legit = 0;
        { std::ifstream f("sobolFnVal.dat"); std::string line; while (std::getline(f, line)) { std::vector<double> v = nums(line, 1); if (!v.empty()) numsolved = std::max(numsolved, (long)v[0]); } }
        { std::ifstream f("legitSobol.dat"); f >> legit; }
        char b[600];
        if (option == 2 && numsolved >= 0) {
            if (numsolved >= c.qr || legit >= c.legit) { /* :189-195 */
                snprintf(b, sizeof b, "Update sobol: numsolved > p_qr_ndraw or legitSobol > p_legitimate %d %d %ld %d Change p_qr_ndraw to %ld or p_legitimate to %ld",
                         oldc.qr, c.qr, legit, c.legit, std::max<long>(numsolved + 1, c.qr), std::max<long>(legit + 1, c.legit));
                return exit_state(b);
            }
        }
        if (option == 3) {
            if (oldc.qr != c.qr) return exit_state("Update LocalMinimizations: old_p_qr_ndraw .NE. p_qr_ndraws. Run with option=2, update sobol points first"); /* :205-211 */
            if (numsolved < 0) return exit_state("Update LocalMinimizations:  sobolFnVal.dat cannot be found!");                                               /* :224-227 */
            if (numsolved < c.maxpoints + 1) return exit_state("Update LocalMinimizations: numsolved < p_maxpoints. Run with option=2, update sobol points first!"); /* :219-223 */
            std::ifstream f("searchResults.dat"); std::string line;
            while (std::getline(f, line)) {
                std::vector<double> v = nums(line, (size_t)c.nx + 4);
                if (v.size() < (size_t)c.nx + 4) continue;
                if ((int)v[1] >= 1) { old_rows.push_back(v); resume_k = std::max(resume_k, (int)v[1]); }
            }
            if (resume_k > c.maxpoints + 1) { /* :238-243 */
                snprintf(b, sizeof b, "Update LocalMinimizations: numsolved > p_maxpoints %d %d Change p_maxpoints to > %d", oldc.maxpoints + 1, c.maxpoints + 1, resume_k + 1);
                return exit_state(b);
            }
            /* gaps (restarts a killed instance never reported) are legal: the run continues after the largest k (genericParams.f90:
             * 229-236) and the missing-search sweep of states 8-9 fills them afterwards (:1048-1146) */
        }
    }
    write_internal_config(c);
    auto envi = [](const char* k, int d) { const char* v = getenv(k); return v ? atoi(v) : d; };
    tiktak_config tc;
    memset(&tc, 0, sizeof tc);
    tc.nx = c.nx; tc.nm = c.nm; tc.maxeval = c.maxeval; tc.qr_ndraw = c.qr; tc.legitimate = c.legit; tc.maxpoints_plus1 = c.maxpoints + 1;
    tc.search_type = c.stype; tc.lottery_points = c.lottery; tc.fvalmax = c.fvalmax;
    tc.range = c.range.data(); tc.init = c.init.data(); tc.bound = c.bound.data();
    tc.objective_id = envi("TIKTAK_OBJECTIVE", 0); tc.device = envi("TIKTAK_DEVICE", 0); tc.rank = 0; tc.world = 1;
    tc.nm_itmax = 5000; tc.nm_ftol = 1.0e-4; tc.text_roundtrip = 1; tc.round_size = envi("TIKTAK_ROUND_SIZE", 1);
    tiktak_handle* h = nullptr;
    const int K = tc.maxpoints_plus1, nx = c.nx, seqNo = 1;
    write_int("state.dat", 0);
    const int ngpu = std::max(1, envi("TIKTAK_NGPU", 1));
    std::vector<tiktak_handle*> hs(ngpu, nullptr);
    for (int r = 0; r < ngpu; r++) { /* rank r of ngpu on device TIKTAK_DEVICE + r */
        tiktak_config tr = tc;
        tr.rank = r; tr.world = ngpu; tr.device = tc.device + r;
        CHECK(tiktak_cuda_create(&hs[r], &tr));
    }
    h = hs[0];
    if (ngpu > 1) CHECK(tiktak_cuda_comm_init_all(hs.data(), ngpu));
    write_int("seqNo.dat", seqNo);
    if (option == 4) { /* :138-149 */
        double f;
        printf(" Running diagnostics for the initial guess:\n");
        CHECK(tiktak_cuda_objfun(h, c.init.data(), 1, &f));
        printf("Objective value = %12.6f\n", f);
        for (int r = 0; r < ngpu; r++) tiktak_cuda_destroy(hs[r]);
        return 0;
    }
    if (option == 3) { for (const char* fn : {"sobolFnVal.dat", "FinalStart.dat", "FinalResults.dat"}) { FILE* f = fopen(fn, "w"); if (f) fclose(f); } }
    else for (const char* fn : {"sobolFnVal.dat", "searchStart.dat", "searchResults.dat", "FinalStart.dat", "FinalResults.dat", "logFile.txt"}) { FILE* f = fopen(fn, "w"); if (f) fclose(f); }
    /* ranks 1.. of a multi-GPU run: the same stage calls (each is collective over the handles), no files */
    std::vector<std::thread> helpers;
    std::vector<int> helper_rc(ngpu, 0);
    std::atomic<bool> values_read{false}; /* rank 0 has copied every shard's values out for sobolFnVal.dat: the handles are free again */
    {
        std::vector<int> helper_missing; /* the same list rank 0 derives below (states 8-9) */
        if (option != 5) {
            std::vector<char> solved(K + 1, 0);
            for (auto& r : old_rows) if ((int)r[1] >= 1 && (int)r[1] <= K) solved[(int)r[1]] = 1;
            for (int k = 1; k <= resume_k; k++) if (!solved[k]) helper_missing.push_back(k);
        }
        const int n_old = (int)old_rows.size();
        std::vector<double> old_cm((size_t)n_old * (nx + 4));
        for (int r = 0; r < n_old; r++) for (int cc = 0; cc < nx + 4; cc++) old_cm[(size_t)cc * n_old + r] = old_rows[r][cc];
        for (int r = 1; r < ngpu; r++)
            helpers.emplace_back([&, r, old_cm, n_old, helper_missing]() {
                tiktak_handle* hr = hs[r];
                int64_t e = 0, l = 0;
                int rc = tiktak_cuda_pretest(hr, &e, &l);
                while (!values_read.load()) std::this_thread::sleep_for(std::chrono::microseconds(200));
                if (!rc) rc = tiktak_cuda_choose(hr, nullptr, nullptr);
                if (!rc && n_old > 0) rc = tiktak_cuda_push_results(hr, old_cm.data(), n_old);
                const int B = tc.round_size < 1 ? 1 : tc.round_size;
                int k0 = resume_k + 1;
                const int kend = option == 5 ? 1 : K;
                int32_t window = 0;
                if (!rc) rc = tiktak_cuda_local_capacity(hr, alg, &window);
                const int per_launch = std::max(B, (window / B) * B);
                while (!rc && k0 <= kend) {
                    int32_t acc = 0;
                    rc = tiktak_cuda_local_launch(hr, alg, k0, std::min(per_launch, kend - k0 + 1), B, &acc);
                    k0 += acc;
                }
                for (size_t q = 0; !rc && q < helper_missing.size(); q++) rc = tiktak_cuda_local_missing(hr, alg, helper_missing[q], nullptr, nullptr, nullptr);
                helper_rc[r] = rc;
            });
    }
    struct Joiner { std::vector<std::thread>& t; ~Joiner() { for (auto& x : t) if (x.joinable()) x.detach(); } } joiner{helpers};
    /* states 2-5 */
    write_int("state.dat", 2);
    if (envi("TIKTAK_WRITE_SOBOL", 0) && c.qr > 0) {
        std::vector<double> pts((size_t)c.qr * nx);
        CHECK(tiktak_cuda_sobol_points(h, 1, c.qr, pts.data()));
        mywrite2("sobol.dat", pts.data(), c.qr, nx);
    }
    write_int("state.dat", 3);
    int64_t ne = 0, nl = 0;
    CHECK(tiktak_cuda_pretest(h, &ne, &nl));
    {
        FILE* f = fopen("sobolFnVal.dat", "a");
        const int64_t chunk = 65536;
        std::vector<double> fv(chunk), pts((size_t)chunk * nx);
        for (int64_t first = 1; first <= ne; first += chunk) {
            const int64_t cnt = std::min<int64_t>(chunk, ne - first + 1);
            CHECK(tiktak_cuda_pretest_values(h, first, cnt, fv.data()));
            if (ngpu > 1) { /* a shard returns NaN for the count blocks of the others: take each value from its owner */
                std::vector<double> part(cnt);
                for (int r = 1; r < ngpu; r++) {
                    CHECK(tiktak_cuda_pretest_values(hs[r], first, cnt, part.data()));
                    for (int64_t q = 0; q < cnt; q++) if (part[q] == part[q]) fv[q] = part[q];
                }
            }
            CHECK(tiktak_cuda_sobol_points(h, first, cnt, pts.data()));
            for (int64_t r = 0; r < cnt; r++) {
                fprintf(f, "%8lld%40.20f", (long long)(first + r), fv[r]);
                for (int d = 0; d < nx; d++) fprintf(f, "%40.20f", pts[(size_t)d * cnt + r]);
                fputc('\n', f);
            }
        }
        fclose(f);
    }
    values_read.store(true);
    write_int("lastSobol.dat", (long)ne + 1);
    write_int("legitSobol.dat", (long)nl);
    /* states 4-5 (setupMissingSobol / solveMissingSobol, :901-1046): the device evaluates the whole prefix 1..Kcut in one stage, so
     * no Sobol point can be missing from sobolFnVal.dat; the reference's bookkeeping files are written for the instances that read them */
    { FILE* f = fopen("missingSobol.dat", "w"); if (f) fclose(f); }
    write_int("legitSobolMiss.dat", (long)nl);
    { char b[200]; snprintf(b, sizeof b, "%d has evaluated %lld sobol points, %lld legitimate", seqNo, (long long)ne, (long long)nl); logmsg(b); }
    /* state 6 */
    write_int("state.dat", 6);
    std::vector<double> xs((size_t)K * (nx + 1));
    std::vector<int32_t> idx(K);
    CHECK(tiktak_cuda_choose(h, xs.data(), idx.data()));
    mywrite2("x_starts.dat", xs.data(), K, nx + 1);
    /* state 7 */
    write_int("state.dat", 7);
    {
        FILE* fr = fopen("searchResults.dat", "a"); FILE* fs = fopen("searchStart.dat", "a");
        if (resume_k == 0) { /* whichPoint == 1: the raw p_init record (:496-501) */
            fprintf(fr, "%10d%10d%10d", seqNo, 0, 0); for (int c2 = 0; c2 <= nx; c2++) fprintf(fr, "%40.20f", xs[(size_t)c2 * K]); fputc('\n', fr);
            fprintf(fs, "%10d%10d", seqNo, 0); for (int c2 = 1; c2 <= nx; c2++) fprintf(fs, "%40.20f", xs[(size_t)c2 * K]); fputc('\n', fs);
        } else { /* option 3: the finished restarts go back into the device-resident results table, in file order */
            const int n = (int)old_rows.size();
            std::vector<double> rows((size_t)n * (nx + 4));
            for (int r = 0; r < n; r++) for (int cc = 0; cc < nx + 4; cc++) rows[(size_t)cc * n + r] = old_rows[r][cc];
            CHECK(tiktak_cuda_push_results(h, rows.data(), n));
            char b[200]; snprintf(b, sizeof b, "%d resumes the local searches after restart %d (%d rows of searchResults.dat loaded)", seqNo, resume_k, n); logmsg(b);
        }
        const int B = tc.round_size < 1 ? 1 : tc.round_size;
        int k0 = resume_k + 1, kend = option == 5 ? 1 : K; /* option 5: only the initial guess (:151-197) */
        /* Rounds of B restarts (B = 1: the single-instance order of the reference).  A launch holds one resident wave
         * of restarts, i.e. several rounds started early against the current best row; the library accepts them
         * while that row stays the best one and the rest is queued again -- same rows as one round after the other. */
        int32_t window = 0;
        CHECK(tiktak_cuda_local_capacity(h, alg, &window));
        const int per_launch = std::max(B, (window / B) * B);
        const int kfirst = k0;
        /* TIKTAK_ASYNC_INSTANCES=<n> (one GPU): the restarts as n asynchronous instances -- the reference's multi-process mode in one
         * kernel launch (tiktak_cuda_local_async) -- instead of rounds.  Rows are written in restart order; their #improvements
         * column counts in the order the instances finished (the order of the file n processes would have written). */
        if (const char* ea = getenv("TIKTAK_ASYNC_INSTANCES")) {
            int32_t yes = 0, used = 0;
            CHECK(tiktak_cuda_local_async_supported(h, alg, &yes));
            if (yes && atoi(ea) > 0 && option != 5 && k0 <= kend) {
                CHECK(tiktak_cuda_local_async(h, alg, k0, kend - k0 + 1, atoi(ea), &used));
                char b[200]; snprintf(b, sizeof b, "%d runs restarts %d..%d as %d asynchronous instances", seqNo, k0, kend, (int)used); logmsg(b);
                k0 = kend + 1;
            }
        }
        while (k0 <= kend) {
            const int nk = std::min(per_launch, kend - k0 + 1);
            int32_t acc = 0;
            CHECK(tiktak_cuda_local_launch(h, alg, k0, nk, B, &acc));
            k0 += acc;
        }
        if (kend >= kfirst) {
            const int nk = kend - kfirst + 1;
            std::vector<double> st((size_t)nk * nx), rs((size_t)nk * (nx + 4));
            std::vector<int32_t> ev(nk);
            CHECK(tiktak_cuda_local_fetch(h, kfirst, nk, st.data(), rs.data(), ev.data()));
            {
                std::vector<int32_t> stt(nk);
                CHECK(tiktak_cuda_local_status(h, kfirst, nk, stt.data()));
                int n1 = 0;
                for (int s = 0; s < nk; s++) n1 += stt[s] != 0;
                if (n1) { char b[200]; snprintf(b, sizeof b, "%d: %d local search(es) stopped early (status != 0: DFNLS refused its start)", seqNo, n1); logmsg(b); }
                CHECK(tiktak_cuda_local_nrescue(h, kfirst, nk, stt.data()));
                int nr = 0;
                for (int s = 0; s < nk; s++) nr += stt[s];
                if (nr) { char b[200]; snprintf(b, sizeof b, "%d: DFNLS rebuilt its interpolation system %d time(s) (RESCUE_H)", seqNo, nr); logmsg(b); }
            }
            for (int s = 0; s < nk; s++) {
                fprintf(fs, "%10d%10d", seqNo, kfirst + s); for (int d = 0; d < nx; d++) fprintf(fs, "%40.20f", st[(size_t)d * nk + s]); fputc('\n', fs);
                fprintf(fr, "%10d%10d%10d", seqNo, kfirst + s, (int)rs[(size_t)2 * nk + s]);
                for (int d = 3; d < nx + 4; d++) fprintf(fr, "%40.20f", rs[(size_t)d * nk + s]);
                fputc('\n', fr);
            }
        }
        /* states 8-9: setupMissingSearch / SolveMissingSearch (:1048-1146) -- restarts missing from searchResults.dat */
        std::vector<int> missing;
        if (option != 5) {
            std::vector<char> solved(K + 1, 0);
            for (auto& r : old_rows) if ((int)r[1] >= 1 && (int)r[1] <= K) solved[(int)r[1]] = 1;
            for (int k = 1; k <= resume_k; k++) if (!solved[k]) missing.push_back(k);
        }
        { FILE* f = fopen("missingSearch.dat", "w"); if (f) { for (int k : missing) fprintf(f, "%10d\n", k); fclose(f); } }
        if (!missing.empty()) {
            write_int("state.dat", 9);
            for (int k : missing) {
                char b[200]; snprintf(b, sizeof b, "%d searching using sobol point %d (missing from searchResults.dat)", seqNo, k); logmsg(b);
                std::vector<double> st(nx), rs(nx + 4); int32_t ev = 0;
                CHECK(tiktak_cuda_local_missing(h, alg, k, st.data(), rs.data(), &ev));
                fprintf(fs, "%10d%10d", seqNo, k); for (int d = 0; d < nx; d++) fprintf(fs, "%40.20f", st[d]); fputc('\n', fs);
                fprintf(fr, "%10d%10d%10d", seqNo, k, (int)rs[2]); for (int d = 3; d < nx + 4; d++) fprintf(fr, "%40.20f", rs[d]); fputc('\n', fr);
            }
            { FILE* f = fopen("missingSearch.dat", "w"); if (f) fclose(f); } /* the list is consumed entry by entry (:1112-1118) */
        }
        fclose(fr); fclose(fs);
        write_int("lastParam.dat", (long)kend + 1);
    }
    /* state 10: lastSearch */
    write_int("state.dat", 10);
    {
        double fb; std::vector<double> xb(nx); int32_t kb, it;
        CHECK(tiktak_cuda_best(h, &fb, xb.data(), &kb));
        FILE* f = fopen("FinalStart.dat", "a");
        fprintf(f, "%10d%10d", seqNo, kb); for (int d = 0; d < nx; d++) fprintf(f, "%40.20f", xb[d]); fputc('\n', f); fclose(f);
        double ff; std::vector<double> xf(nx);
        CHECK(tiktak_cuda_last_search(h, &ff, xf.data(), &it));
        f = fopen("FinalResults.dat", "a");
        fprintf(f, "%10d%10d%40.20f", seqNo, kb, ff); for (int d = 0; d < nx; d++) fprintf(f, "%40.20f", xf[d]); fputc('\n', f); fclose(f);
        char b[300]; snprintf(b, sizeof b, "The best final point is %.12f (restart %d), after the final search %.12f", fb, kb, ff); logmsg(b);
    }
    { char b[100]; snprintf(b, sizeof b, "%d has no more tasks.", seqNo); logmsg(b); }
    for (auto& t : helpers) if (t.joinable()) t.join();
    for (int r = 1; r < ngpu; r++) if (helper_rc[r]) { char b[200]; snprintf(b, sizeof b, "EXIT STATE: rank %d: %s", r, tiktak_cuda_strerror(helper_rc[r])); logmsg(b); return 1; }
    for (int r = 0; r < ngpu; r++) tiktak_cuda_destroy(hs[r]);
    return 0;
}
