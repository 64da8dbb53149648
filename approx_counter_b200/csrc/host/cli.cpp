// cli.cpp — the drop-in `approx_counter` command line: same flags, defaults,
// banner, output files and exit behaviour as the reference's main()
// (/root/reference/approx_counter.cpp:604-669 option table, :679-958 driver),
// with the two hot functions replaced by libapc's C ABI:
//   count_kmers + get_most_frequent (:874, :898) -> apc_exact_topn
//   errorCount                      (:922)       -> apc_approx_count
//   readRecords + sampleSequences   (:825, :867) -> apc_ingest_fastx + apc_sample_resident with --ingest device
// Extensions (not in the reference): --seed, --gpus, --device, --ingest.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "apc.h"
#include "host_util.h"

namespace apch {

namespace {

const auto boot_time = std::chrono::steady_clock::now(); // :19

template <typename T>
void print(const T &text, int tab = 0) { // :85-94
    const auto milis =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - boot_time).count();
    std::cout << "[" << milis << " ms]\t";
    for (int i = 0; i < tab; i++) std::cout << "\t";
    std::cout << text << std::endl;
}

// read-only mapping of the input, paged in by the kernel while mmap runs (MAP_POPULATE) — --ingest device
struct MappedFile {
    const uint8_t *data = nullptr;
    size_t size = 0;
    bool open(const std::string &path) {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
            ::close(fd);
            return false;
        }
        size = (size_t)st.st_size;
        if (size) {
            void *m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
            if (m == MAP_FAILED) {
                ::close(fd);
                return false;
            }
            data = (const uint8_t *)m;
        }
        ::close(fd);
        return true;
    }
    ~MappedFile() {
        if (data) munmap((void *)data, size);
    }
};

enum class Kind { Int, Double, String, Flag };
struct OptSpec {
    const char *short_name, *long_name;
    Kind kind;
    const char *help;
};
// option table of get_args (:606-662) + extensions
const OptSpec OPTIONS[] = {
    {"lc", "low_complexity", Kind::Double, "low complexity filter threshold (for k=16), default 1.5"},
    {"sn", "sample_n", Kind::Int, "sample n sequences from dataset, default 10k sequences"},
    {"sl", "sample_length", Kind::Int, "size of the sampled portion, default 100 bases"},
    {"nt", "nb_thread", Kind::Int, "Number of thread to work with, default is 4 (ignored: the count runs on the GPU)"},
    {"k", "kmer_size", Kind::Int, "Size of the kmers, default is 16"},
    {"lim", "limit", Kind::Int, "limit the number of kmer used after initial counting, default is 500"},
    {"mr", "multi_run", Kind::Int, "Number of time the count must be performed. Each count is exported separately."},
    {"v", "verbosity", Kind::Int, "Level of details printed out"},
    {"e", "exact_file", Kind::String, "path to export the exact k-mer count, if needed. Default: no export"},
    {"conf", "config", Kind::String, "path to the config file"},
    {"fk", "forbidden_kmer", Kind::String,
     "take a file containing 'forbidden' kmers, excluding them from the search pool. One kmer per line."},
    {"sk", "solid_km", Kind::Int,
     "Use solid kmer instead of most frequents. This option will override sample number (-sn / --sample_n)."},
    {"se", "skip_end", Kind::Flag, "Skip end adapter ressearch (only search start)."},
    {"o", "out_file", Kind::String, "path to the output file, default is ./out.txt"},
    {"", "seed", Kind::Int, "[extension] seed of the read shuffle (default: std::random_device, like the reference)"},
    {"", "gpus", Kind::Int, "[extension] number of GPUs to shard the sampled reads over (default 1)"},
    {"", "device", Kind::Int, "[extension] first CUDA device to use (default 0)"},
    {"", "ingest", Kind::String,
     "[extension] device (default): the input's bytes are copied to the GPU, indexed and sampled there (FASTA, 4-line "
     "FASTQ; wrapped FASTQ and blanks inside sequence lines are handed to the host parser; with --gpus N the other GPUs "
     "fetch their shard of the sample from the first, GPU to GPU); host: parsed and sampled by host threads"},
    {"", "version-check", Kind::String, "[accepted for SeqAn compatibility, ignored]"},
};

struct Parsed {
    std::map<std::string, std::string> values; // by long name
    std::vector<std::string> positional;
    bool isSet(const char *long_name) const { return values.count(long_name) > 0; }
};

enum ParseResult { PARSE_OK, PARSE_ERROR, PARSE_HELP };

void print_help() {
    std::cout << "adaptFinder\n===========\n\nSYNOPSIS\n    approx_counter [OPTIONS] <input filename>\n\nOPTIONS\n";
    for (const auto &o : OPTIONS) {
        std::cout << "    ";
        if (o.short_name[0]) std::cout << "-" << o.short_name << ", ";
        std::cout << "--" << o.long_name;
        if (o.kind != Kind::Flag) std::cout << (o.kind == Kind::Int ? " INT" : o.kind == Kind::Double ? " DOUBLE" : " STRING");
        std::cout << "\n          " << o.help << "\n";
    }
}

bool valid_int(const std::string &s) {
    if (s.empty()) return false;
    size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (i == s.size()) return false;
    for (; i < s.size(); i++)
        if (s[i] < '0' || s[i] > '9') return false;
    return true;
}
bool valid_double(const std::string &s) {
    if (s.empty()) return false;
    char *end = nullptr;
    std::strtod(s.c_str(), &end);
    return end && *end == '\0';
}

ParseResult parse_args(int argc, const char **argv, Parsed &out) {
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-h" || a == "--help") { print_help(); return PARSE_HELP; }
        if (a == "--version") { std::cout << "approx_counter (B200) libapc " << apc_version() << "\n"; return PARSE_HELP; }
        if (a.size() >= 2 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) {
            const bool is_long = a[1] == '-';
            std::string name = a.substr(is_long ? 2 : 1), inline_val;
            bool has_inline = false;
            const size_t eq = name.find('=');
            if (eq != std::string::npos) { inline_val = name.substr(eq + 1); name = name.substr(0, eq); has_inline = true; }
            const OptSpec *spec = nullptr;
            for (const auto &o : OPTIONS)
                if (name == (is_long ? o.long_name : o.short_name) && name[0]) spec = &o;
            if (!spec) {
                std::cerr << "approx_counter: Unknown option: " << a << "\n";
                return PARSE_ERROR;
            }
            if (spec->kind == Kind::Flag) { out.values[spec->long_name] = "1"; continue; }
            std::string val;
            if (has_inline) val = inline_val;
            else if (i + 1 < argc) val = argv[++i];
            else {
                std::cerr << "approx_counter: option requires an argument: " << a << "\n";
                return PARSE_ERROR;
            }
            if ((spec->kind == Kind::Int && !valid_int(val)) || (spec->kind == Kind::Double && !valid_double(val))) {
                std::cerr << "approx_counter: the given value '" << val << "' cannot be casted to "
                          << (spec->kind == Kind::Int ? "integer" : "double") << " for option " << a << "\n";
                return PARSE_ERROR;
            }
            out.values[spec->long_name] = val;
        } else {
            out.positional.push_back(a);
        }
    }
    if (out.positional.size() < 1) { std::cerr << "approx_counter: Not enough arguments were provided.\n"; return PARSE_ERROR; }
    if (out.positional.size() > 1) { std::cerr << "approx_counter: Too many arguments were provided.\n"; return PARSE_ERROR; }
    return PARSE_OK;
}

void get_u64(const Parsed &p, const char *name, uint64_t &v) { // getOptionValue: untouched if not set
    auto it = p.values.find(name);
    if (it != p.values.end()) v = (uint64_t)std::strtoll(it->second.c_str(), nullptr, 10);
}
void get_str(const Parsed &p, const char *name, std::string &v) {
    auto it = p.values.find(name);
    if (it != p.values.end()) v = it->second;
}

struct Gpu {
    apc_ctx *ctx = nullptr;
    ~Gpu() { apc_destroy(ctx); }
};

int gpu_fail(const char *what, apc_ctx *ctx, int st) {
    std::cerr << "/!\\ ERROR: " << what << ": " << apc_strerror(st);
    if (ctx && apc_last_error(ctx)[0]) std::cerr << " (" << apc_last_error(ctx) << ")";
    std::cerr << std::endl;
    return 2;
}

} // namespace

bool one_shot_process = false; // set by main.cpp

int cli_main(int argc, const char **argv) {
    Parsed parser;
    const ParseResult res = parse_args(argc, argv, parser);
    if (res != PARSE_OK) return res == PARSE_ERROR; // :697-698

    // Default values for parameters (:701-715)
    std::string output = "out.txt", exact_out, config_file, forbid_kmer;
    uint64_t solid_km = 0, nb_thread = 4, k = 16, sl = 100, sn = 40000, limit = 500;
    float param_lc = 1.0;
    uint64_t v = 1;
    bool skip_end = false;
    uint64_t nb_of_runs = 1;
    float lc = 1.0;
    uint64_t n_gpus = 1, device0 = 0;
    int64_t seed = -1;
    std::string ingest = "device";

    get_str(parser, "config", config_file);
    if (!config_file.empty()) { // :721-737
        std::vector<std::pair<std::string, std::string>> kv;
        parse_config(config_file, kv);
        std::map<std::string, std::string> params;
        for (auto &p : kv) params[p.first] = p.second;
        auto has = [&](const char *key) { return params.count(key) > 0; };
        param_lc = has("lc") ? std::stof(params["lc"]) : lc;
        k = has("k") ? std::stoi(params["k"]) : k;
        v = has("v") ? std::stoi(params["v"]) : v;
        sn = has("sn") ? std::stoi(params["sn"]) : sn;
        sl = has("sl") ? std::stoi(params["sl"]) : sl;
        limit = has("lim") ? std::stoi(params["lim"]) : limit;
        nb_thread = has("nt") ? std::stoi(params["nt"]) : nb_thread;
        solid_km = has("sk") ? std::stoi(params["sk"]) : solid_km;
        skip_end = has("se");
        forbid_kmer = has("fk") ? params["fk"] : forbid_kmer;
        exact_out = has("e") ? params["e"] : exact_out;
        nb_of_runs = has("mr") ? std::stoi(params["mr"]) : nb_of_runs;
    }

    // command line overrides config (:744-755)
    get_u64(parser, "limit", limit);
    if (parser.isSet("low_complexity")) param_lc = std::strtof(parser.values["low_complexity"].c_str(), nullptr);
    get_u64(parser, "kmer_size", k);
    get_u64(parser, "verbosity", v);
    get_u64(parser, "sample_length", sl);
    get_u64(parser, "sample_n", sn);
    get_u64(parser, "nb_thread", nb_thread);
    get_str(parser, "out_file", output);
    get_str(parser, "exact_file", exact_out);
    get_str(parser, "forbidden_kmer", forbid_kmer);
    get_u64(parser, "solid_km", solid_km);
    get_u64(parser, "multi_run", nb_of_runs);
    get_u64(parser, "gpus", n_gpus);
    get_u64(parser, "device", device0);
    get_str(parser, "ingest", ingest);
    if (parser.isSet("seed")) seed = std::strtoll(parser.values["seed"].c_str(), nullptr, 10);
    skip_end = skip_end || parser.isSet("skip_end"); // :758

    const std::string input_file = parser.positional[0];

    std::vector<uint64_t> kmer_set; // forbidden k-mers (:764-769)
    if (!forbid_kmer.empty()) {
        print("Parsing the fobidden kmer list");
        if (!parse_kmer_list(forbid_kmer, kmer_set)) {
            std::cerr << "/!\\ ERROR: COULD NOT OPEN EXCLUDED KMER FILE, must quit\n";
            exit(1); // :361
        }
    }

    int mr_v = (int)v; // :772-775
    if (nb_of_runs > 1 && v < 2) mr_v = 0;

    const std::string warning = "/!\\ WARNING: ", error_pref = "/!\\ ERROR: ";

    // uncaught, as in the reference (:781-787): the process aborts
    if (k < 2 || k > 32) throw std::invalid_argument(error_pref + "kmer size must be between 2 and 32 (included)");
    if (k > sl) throw std::invalid_argument(error_pref + "kmer size must be smaller than the sampling length (k <= sl)");

    lc = adjust_threshold(param_lc, 16, (uint8_t)k); // :790

    if (v > 0) { // :793-808
        std::cout << "Kmer size:             " << k << std::endl;
        std::cout << "Sampled sequences:     " << sn << std::endl;
        std::cout << "Sampling length        " << sl << std::endl;
        std::cout << "LC filter threshold:   " << param_lc << std::endl;
        std::cout << "Adjusted LC threshold: " << lc << std::endl;
        std::cout << "Nb thread:             " << nb_thread << std::endl;
        if (solid_km != 0) std::cout << "Solid kmers:           " << solid_km << std::endl;
        else std::cout << "Number of kept kmer:   " << limit << std::endl;
        std::cout << "Number of runs:        " << nb_of_runs << std::endl;
        std::cout << "Verbosity level:       " << v << std::endl;
    }

    int tab_level = 0;
    if (v > 0 && nb_of_runs > 1) std::cout << "\nA total of " << nb_of_runs << " runs will be performed." << std::endl;

    // GPU contexts: no CPU fallback — fail loudly if the device is unusable.  Creating a CUDA
    // context takes about half a second, so it runs beside the parsing of the input file.
    if (n_gpus < 1) n_gpus = 1;
    std::vector<Gpu> gpus(n_gpus);
    int create_status = APC_OK;
    double create_ms = 0., reserve_ms = 0.;
    std::thread creator([&]() {
        const auto t0 = std::chrono::steady_clock::now();
        for (uint64_t g = 0; g < n_gpus && create_status == APC_OK; g++)
            create_status = apc_create((int)(device0 + g), &gpus[g].ctx);
        // buffers for the largest sample this run can draw (sn reads of sl + 1 bases) and every kernel of this k,
        // while the main thread parses the input; a failure here is not fatal, the calls below allocate on demand
        create_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (create_status == APC_OK && k <= sl)
            apc_reserve(gpus[0].ctx, std::min<uint64_t>(sn, 4u << 20), (uint32_t)(sl + 1), (uint8_t)k, (uint32_t)std::min<uint64_t>(limit, 1u << 20));
        reserve_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() - create_ms;
    });
    struct Joiner { // an exception out of read_fastx (bad_alloc on a huge input) must not leave the thread joinable
        std::thread &t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } creator_guard{creator};

    if (ingest != "host" && ingest != "device") {
        std::cerr << error_pref << "--ingest takes host or device" << std::endl;
        return 1;
    }
    if (v > 0) print("Parsing FASTA file", tab_level); // :821-825
    Reads seqs;
    uint64_t n_seqs = 0;
    bool device_ingest = false;
    if (ingest == "device") {
        // The file is mapped and paged in beside the creation of the CUDA context, then copied to the GPU, where the
        // records are indexed and every sample is gathered (apc_ingest_fastx / apc_sample_resident).
        MappedFile file;
        if (!file.open(input_file)) {
            // a pipe, or a file that is not there: the host parser reads the former and reports the latter
            if (v > 1) print("Input cannot be mapped; using the host parser", tab_level);
        } else {
        if (v > 1) print("File mapped; waiting for the CUDA context", tab_level);
        creator.join();
        if (create_status != APC_OK) return gpu_fail("cannot open CUDA device", nullptr, create_status);
        if (v > 1)
            print("CUDA context ready (beside the mapping: context " + std::to_string(create_ms) + " ms, buffers and kernels " +
                      std::to_string(reserve_ms) + " ms); copying the input to the GPU", tab_level);
        int is_fastq = 0;
        const int st = apc_ingest_fastx(gpus[0].ctx, file.data, file.size, &n_seqs, &is_fastq);
        if (st == APC_OK) {
            device_ingest = true;
            if (v > 1) {
                float copy_ms = 0.f, index_ms = 0.f;
                apc_ingest_timing(gpus[0].ctx, &copy_ms, &index_ms, nullptr);
                print(std::string(is_fastq ? "FASTQ" : "FASTA") + " indexed on the GPU (copy " + std::to_string(copy_ms) +
                          " ms, index " + std::to_string(index_ms) + " ms)", tab_level);
            }
        } else if (st == APC_ERR_FORMAT) {
            if (v > 1) print(std::string("Device parser: ") + apc_last_error(gpus[0].ctx) + "; using the host parser", tab_level);
        } else {
            return gpu_fail("copying the input to the GPU", gpus[0].ctx, st);
        }
        }
    }
    if (!device_ingest) {
        std::string err;
        const bool parsed = read_fastx(input_file, seqs, err);
        n_seqs = seqs.size();
        if (v > 1) print("File parsed; waiting for the CUDA context", tab_level);
        if (creator.joinable()) creator.join();
        if (v > 1)
            print("CUDA context ready (beside the parsing: context " + std::to_string(create_ms) + " ms, buffers and kernels " +
                      std::to_string(reserve_ms) + " ms)", tab_level);
        if (create_status != APC_OK) return gpu_fail("cannot open CUDA device", nullptr, create_status);
        if (!parsed) {
            std::cerr << error_pref << err << std::endl;
            return 1;
        }
    }
    apc_ctx *ctx0 = gpus[0].ctx;
    // several GPUs: the contexts form one NCCL communicator, one rank each (apc_comm_*, include/apc.h)
    bool have_comm = false;
    if (n_gpus > 1) {
        uint8_t id[APC_COMM_ID_BYTES];
        int st = apc_comm_unique_id(id);
        if (st == APC_OK) {
            std::vector<int> status(n_gpus, APC_OK);
            std::vector<std::thread> joiners;
            for (uint64_t g = 0; g < n_gpus; g++)
                joiners.emplace_back([&, g]() { status[g] = apc_comm_init_rank(gpus[g].ctx, (int)n_gpus, (int)g, id); });
            for (auto &t : joiners) t.join();
            have_comm = true;
            for (uint64_t g = 0; g < n_gpus; g++)
                if (status[g] != APC_OK) { have_comm = false; st = status[g]; }
        }
        if (!have_comm) {
            std::cerr << warning << "no NCCL communicator (" << apc_strerror(st) << ": " << apc_last_error(ctx0)
                      << "); the per-GPU count vectors are summed on the host" << std::endl;
            for (uint64_t g = 0; g < n_gpus; g++) apc_comm_destroy(gpus[g].ctx);
        } else if (v > 1) {
            print("NCCL communicator over " + std::to_string(n_gpus) + " GPUs", tab_level);
        }
    }
    if (v > 0) print("Number of sequences found: " + std::to_string(n_seqs) + ".", tab_level);

    std::string run_suffix;
    for (uint64_t current_run = 0; current_run < nb_of_runs; current_run++) { // :835
        run_suffix = "_" + std::to_string(current_run);                       // :837 (always appended)
        if (nb_of_runs > 1 && v > 0) std::cout << "Starting run number " << current_run + 1 << std::endl;
        const uint64_t sequence_set_size = n_seqs;
        if (sn > sequence_set_size) { // :844-848
            std::cerr << warning << "Sequence set too small for the requested sample size\n";
            std::cerr << warning << "The whole set will be used.\n";
            sn = sequence_set_size;
        }
        bool success = true;
        const std::array<std::string, 2> ends = {"start", "end"};
        bool bottom = false;
        tab_level += 1;
        for (const std::string &which_end : ends) { // :858
            if (v > 0) print("Working on sequence " + which_end + ".", tab_level - 1);
            if (mr_v > 0) print("Sampling", tab_level);
            if (mr_v > 0) print(bottom ? "Sampling the ends of reads" : "Sampling the start of reads", 1);
            uint64_t n_sampled = 0;
            uint32_t row_len = 0;
            SampleBytes sample;
            int st = APC_OK;
            if (device_ingest) { // :867 on the device: the host only shuffles the ids (:423-429)
                // when every read is wanted (sn >= #reads, :844-848) the walk takes ALL eligible reads whatever the
                // order, and nothing downstream depends on the order of the sample's rows (counts are sums over reads):
                // the shuffle (6 ms per million ids) is skipped and the file order used
                std::vector<int> order;
                if (sn < n_seqs) order = shuffle_order(n_seqs, seed);
                st = apc_sample_resident(ctx0, order.empty() ? nullptr : reinterpret_cast<const uint32_t *>(order.data()),
                                         order.size(), sn, (uint32_t)sl, bottom ? 1 : 0, &n_sampled);
                if (st != APC_OK) return gpu_fail("sampling on the GPU", ctx0, st);
                row_len = (uint32_t)(sl + (bottom ? 1 : 0));
            } else {
                sample = sample_sequences(seqs, sn, sl, bottom, seed, n_sampled, row_len); // :867
            }
            if (mr_v > 0) print("Sampled " + std::to_string(n_sampled) + " sequences", 1);

            if (!device_ingest) st = apc_upload_sample(ctx0, sample.data(), n_sampled, row_len);
            if (st != APC_OK) return gpu_fail("uploading the sample", ctx0, st);

            if (mr_v > 0) print("Exact k-mer count", tab_level); // :872-874
            std::vector<uint64_t> km, ct;
            uint64_t n_top = 0, n_distinct = 0, had_n = 0;
            if (solid_km != 0) {
                uint64_t cap = 1 << 16;
                for (;;) {
                    km.assign(cap, 0);
                    ct.assign(cap, 0);
                    st = apc_exact_solid(ctx0, (uint8_t)k, lc, solid_km, kmer_set.data(), kmer_set.size(), km.data(),
                                         ct.data(), cap, &n_top, &n_distinct, &had_n);
                    if (st == APC_ERR_CAPACITY) { cap = n_top; continue; }
                    break;
                }
            } else {
                km.assign(limit ? limit : 1, 0);
                ct.assign(limit ? limit : 1, 0);
                st = apc_exact_topn(ctx0, (uint8_t)k, lc, limit, kmer_set.data(), kmer_set.size(), km.data(), ct.data(),
                                    &n_top, &n_distinct, &had_n);
            }
            if (st != APC_OK) return gpu_fail("exact k-mer count", ctx0, st);
            if (had_n > 0) { // :513-517
                std::cerr << "/!\\ WARNING: This dataset contained sequences with 'N' symbols. ";
                std::cerr << "/!\\ WARNING: Current implementation ignores k-mers containing 'N'.";
                std::cerr << "/!\\ WARNING: A total of " << had_n << " k-mers were ignored." << std::endl;
            }
            if (mr_v > 0) print("Number of kmer found: " + std::to_string(n_distinct), tab_level);
            if (mr_v > 0) print(solid_km != 0 ? "Keeping solid k-mer" : "Keeping most frequent k-mer", tab_level);
            pair_vector first_n_vector(n_top);
            for (uint64_t i = 0; i < n_top; i++) first_n_vector[i] = {km[i], ct[i]};
            if (mr_v > 0) print("Number of kmer kept:  " + std::to_string(first_n_vector.size()), tab_level);

            if (!exact_out.empty()) { // :907-916
                if (mr_v > 0) print("Exporting exact kmer count", tab_level);
                success = export_counter(first_n_vector, (uint8_t)k, exact_out + run_suffix + "." + which_end);
                if (!success) {
                    std::cerr << error_pref + "Failed to export exact k-mer count" << std::endl;
                    std::cerr << "Path: " << exact_out + run_suffix + "." + which_end << std::endl;
                    return 1;
                }
            }

            if (mr_v > 0) print("Approximate k-mer count", tab_level); // :919-923
            km.resize(n_top);
            std::vector<uint64_t> approx(n_top, 0);
            if (n_gpus == 1) {
                st = apc_approx_count(ctx0, (uint8_t)k, km.data(), (uint32_t)n_top, approx.data());
                if (st != APC_OK) return gpu_fail("approximate k-mer count", ctx0, st);
            } else {
                // Shard the sampled reads over the GPUs (contiguous blocks of 32-read tiles), one host thread per
                // GPU: upload the shard (GPU 0 already holds the whole sample for the exact stage and scans a
                // sub-range of it), scan it for all k-mers, sum the count vectors with one ncclAllReduce issued
                // through the C ABI (apc_scan_allreduce).  Without a communicator the vectors are summed here.
                const uint64_t tiles = (n_sampled + 31) / 32, per = (tiles + n_gpus - 1) / n_gpus;
                std::vector<int> status(n_gpus, APC_OK);
                std::vector<std::vector<uint64_t>> part(n_gpus);
                std::vector<std::thread> workers;
                for (uint64_t g = 0; g < n_gpus; g++)
                    workers.emplace_back([&, g]() {
                        const uint64_t first = std::min(n_sampled, g * per * 32), last = std::min(n_sampled, (g + 1) * per * 32);
                        apc_ctx *c = gpus[g].ctx;
                        int s = APC_OK;
                        if (g == 0) {
                            apc_set_option(c, "scan_first_read", (int64_t)first);
                            apc_set_option(c, "scan_n_reads", (int64_t)(last - first));
                        } else if (device_ingest) { // the shard comes from GPU 0's resident sample, GPU to GPU
                            s = apc_upload_sample_peer(c, ctx0, first, last - first);
                        } else {
                            s = apc_upload_sample_async(c, sample.data() + first * row_len, last - first, row_len);
                        }
                        if (s == APC_OK) s = apc_set_queries(c, (uint8_t)k, km.data(), (uint32_t)n_top);
                        if (s == APC_OK) s = have_comm ? apc_scan_allreduce(c, nullptr) : apc_scan(c, nullptr);
                        if (s == APC_OK && (g == 0 || !have_comm)) {
                            part[g].assign(n_top, 0);
                            s = apc_get_counts(c, part[g].data());
                        } else if (s == APC_OK) {
                            s = apc_sync(c);
                        }
                        status[g] = s;
                    });
                for (auto &t : workers) t.join();
                for (uint64_t g = 0; g < n_gpus; g++)
                    if (status[g] != APC_OK) return gpu_fail("approximate k-mer count", gpus[g].ctx, status[g]);
                for (uint64_t g = 0; g < (have_comm ? 1 : n_gpus); g++)
                    for (uint64_t i = 0; i < n_top; i++) approx[i] += part[g][i];
                apc_set_option(ctx0, "scan_first_read", 0);
                apc_set_option(ctx0, "scan_n_reads", -1);
            }
            pair_vector sorted_error_count(n_top); // results[kmer] = total (:596), then :923
            for (uint64_t i = 0; i < n_top; i++) sorted_error_count[i] = {km[i], approx[i]};
            get_most_frequent(sorted_error_count, limit, (int)k);

            if (mr_v > 0) print("Exporting approximate count", tab_level);
            success = export_counter(sorted_error_count, (uint8_t)k, output + run_suffix + "." + which_end); // :928
            if (!success) {
                std::cerr << error_pref + "Failed to export approximate k-mer count" << std::endl;
                std::cerr << "Path: " << output + run_suffix + "." + which_end << std::endl;
                return 1;
            }
            if (mr_v > 0) print("Done", tab_level);

            if (skip_end) { // :943-951 — including the reference's quirk: the break only fires when mr_v > 0
                if (mr_v > 0) {
                    print("Skipping end adapter ressearch");
                    break;
                }
            } else {
                bottom = true;
            }
        }
        tab_level--;
    }
    if (one_shot_process) {
        // the binary ends the process right after this call (main.cpp): the driver reclaims the device memory with the
        // context, freeing buffer by buffer first only adds to the wall clock (0.01-1.1 s measured on C3)
        for (Gpu &g : gpus) {
            apc_sync(g.ctx);
            g.ctx = nullptr;
        }
    }
    if (v > 1) print("Releasing the GPU", tab_level);
    gpus.clear();
    if (v > 1) print("Exit", tab_level);
    return 0;
}

} // namespace apch
