// main.cpp -- `scrubby` CLI host: same subcommands and flags as the reference's clap definition
// (terminal.rs:8-50 App/Commands, :206-279 ClassifierArgs, :323-391 AlignmentArgs, :435-466 DiffArgs).
// `reads` (terminal.rs:57-157) runs the external aligner / classifier through `sh -c` with the reference's command
// strings (cleaner.rs:255-649): kraken2 / metabuli outputs and minigraph's stdout PAF come back into the
// GPU path; the SAM-emitting aligners are piped into samtools exactly as the reference does.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>

#include "scrubby_host.hpp"

using namespace scrubby;

namespace {

struct Flag {
    char short_name;
    const char *long_name;
    int arity;  // 0 = switch, 1 = one value, 2 = zero or more values (num_args(0..))
    bool hyphen_ok = false;  // allow_hyphen_values: the value is taken verbatim even when it starts with '-'
};

struct Parsed {
    std::map<std::string, std::vector<std::string>> values;
    std::map<std::string, bool> seen;
    bool has(const std::string &k) const { return seen.count(k) != 0; }
    std::optional<std::string> one(const std::string &k) const {
        auto it = values.find(k);
        if (it == values.end() || it->second.empty()) return std::nullopt;
        return it->second.back();
    }
    std::vector<std::string> many(const std::string &k) const {
        auto it = values.find(k);
        return it == values.end() ? std::vector<std::string>{} : it->second;
    }
};

// `--help` / `-h` (clap prints the help and exits 0): the flags of terminal.rs, per subcommand
const char *HELP_TOP =
    "scrubby 1.0.2 (B200 host)\nRemove or extract background reads (host depletion) from FASTQ / FASTA files\n\n"
    "Usage: scrubby [--log-file <FILE>] <COMMAND>\n\nCommands:\n"
    "  reads       Deplete or extract reads using aligners or classifiers (runs the external tool)\n"
    "  classifier  Deplete or extract reads from classifier outputs (Kraken2 / Metabuli style)\n"
    "  alignment   Deplete or extract reads from an alignment (.paf / .gaf / .sam / .bam) or a read-id list (.txt)\n"
    "  diff        Difference between input and output read files, with optional JSON summary and read-id TSV\n\n"
    "Options:\n  -l, --log-file <FILE>  Output logs to file instead of the terminal\n  -h, --help     Print help\n"
    "  -V, --version  Print version\n\nEnvironment: SCRUBBY_GPU_DEVICE=<index> selects the GPU (default 0)\n";
const char *help_of(const std::string &cmd) {
    if (cmd == "reads")
        return "Usage: scrubby reads [OPTIONS] --index <INDEX>\n\nOptions:\n"
               "  -i, --input [<INPUT>...]            Input read files (one, or two for paired-end; .gz accepted)\n"
               "  -o, --output [<OUTPUT>...]          Output read files (.gz by extension)\n"
               "  -I, --index <INDEX>                 Aligner index file or classifier database directory\n"
               "  -a, --aligner <ALIGNER>             [possible values: bowtie2, minimap2, minigraph, strobealign]\n"
               "  -p, --preset <PRESET>               minimap2 / minigraph preset [lr-hq, splice, splice-hq, asm, asm5, asm10, asm20, sr, lr,\n"
               "                                      map-pb, map-hifi, map-ont, ava-pb, ava-ont]\n"
               "  -c, --classifier <CLASSIFIER>       [possible values: kraken2, metabuli]\n"
               "  -T, --taxa [<TAXA>...]              Taxa and all sub-taxa to deplete (names or taxids)\n"
               "  -D, --taxa-direct [<TAXA_DIRECT>...]  Taxa to deplete directly (names or taxids)\n"
               "  -A, --aligner-args <ARGS>           Additional aligner arguments\n"
               "  -C, --classifier-args <ARGS>        Additional classifier arguments\n"
               "  -t, --threads <THREADS>             Threads for the external tool [default: 4]\n"
               "  -j, --json <JSON>                   Summary report (JSON)\n  -w, --workdir <WORKDIR>             Working directory\n"
               "  -r, --read-ids <READ_IDS>           Read identifiers of depleted / extracted reads (TSV)\n"
               "  -e, --extract                       Extract instead of deplete\n  -h, --help                          Print help\n";
    if (cmd == "classifier")
        return "Usage: scrubby classifier [OPTIONS] --report <REPORT> --reads <READS> --classifier <CLASSIFIER>\n\nOptions:\n"
               "  -i, --input [<INPUT>...]     Input read files\n  -o, --output [<OUTPUT>...]   Output read files\n"
               "  -k, --report <REPORT>        Kraken-style report of the classifier\n"
               "      --reads <READS>          Kraken-style per-read classifications\n"
               "  -c, --classifier <CLASSIFIER>  Output style [possible values: kraken2, metabuli]\n"
               "  -T, --taxa [<TAXA>...]       Taxa and all sub-taxa to deplete\n  -D, --taxa-direct [<TAXA_DIRECT>...]  Taxa to deplete directly\n"
               "  -j, --json <JSON>            Summary report (JSON)\n  -w, --workdir <WORKDIR>      Working directory\n"
               "  -r, --read-ids <READ_IDS>    Read identifiers (TSV)\n  -e, --extract                Extract instead of deplete\n"
               "  -h, --help                   Print help\n";
    if (cmd == "alignment")
        return "Usage: scrubby alignment [OPTIONS] --alignment <ALIGNMENT>\n\nOptions:\n"
               "  -i, --input [<INPUT>...]     Input read files\n  -o, --output [<OUTPUT>...]   Output read files\n"
               "  -a, --alignment <ALIGNMENT>  Alignment (.paf, .gaf, .sam, .bam) or read-id list (.txt)\n"
               "  -f, --format <FORMAT>        Explicit format [possible values: sam, bam, cram, paf, txt, gaf]\n"
               "  -l, --min-len <MIN_LEN>      Minimum query alignment length [default: 0]\n"
               "  -c, --min-cov <MIN_COV>      Minimum query alignment coverage [default: 0]\n"
               "  -q, --min-mapq <MIN_MAPQ>    Minimum mapping quality [default: 0]\n"
               "  -j, --json <JSON>            Summary report (JSON)\n  -w, --workdir <WORKDIR>      Working directory\n"
               "  -r, --read-ids <READ_IDS>    Read identifiers (TSV)\n  -e, --extract                Extract instead of deplete\n"
               "  -h, --help                   Print help\n";
    if (cmd == "diff")
        return "Usage: scrubby diff [OPTIONS]\n\nOptions:\n  -i, --input [<INPUT>...]    Input read files\n"
               "  -o, --output [<OUTPUT>...]  Output read files of a previous run\n  -j, --json <JSON>           Counts (JSON)\n"
               "  -r, --read-ids <READ_IDS>   Read identifiers missing from the output (TSV)\n  -h, --help                  Print help\n";
    return nullptr;
}

[[noreturn]] void usage_error(const std::string &msg) {
    fprintf(stderr, "error: %s\n\nUsage: scrubby [--log-file <FILE>] <reads|classifier|alignment|diff> [OPTIONS]\n", msg.c_str());
    exit(2);
}

Parsed parse(int argc, char **argv, int from, const std::vector<Flag> &flags) {
    Parsed p;
    const Flag *cur = nullptr;
    for (int i = from; i < argc; i++) {
        std::string a = argv[i];
        const Flag *f = nullptr;
        std::string inline_val;
        bool has_inline = false;
        if (cur && cur->arity == 1 && cur->hyphen_ok) {
            p.values[cur->long_name].push_back(a);
            cur = nullptr;
            continue;
        }
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            std::string name = a.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) {
                inline_val = name.substr(eq + 1);
                name = name.substr(0, eq);
                has_inline = true;
            }
            for (auto &fl : flags)
                if (name == fl.long_name) f = &fl;
            if (!f) usage_error("unexpected argument '" + a + "'");
        } else if (a.size() == 2 && a[0] == '-' && a[1] != '-') {
            for (auto &fl : flags)
                if (fl.short_name == a[1]) f = &fl;  // on a short-flag collision the later definition wins
            if (!f) usage_error("unexpected argument '" + a + "'");
        }
        if (f) {
            p.seen[f->long_name] = true;
            cur = f->arity ? f : nullptr;
            if (has_inline) {
                p.values[f->long_name].push_back(inline_val);
                if (f->arity == 1) cur = nullptr;
            }
            continue;
        }
        if (!cur) usage_error("unexpected value '" + a + "'");
        p.values[cur->long_name].push_back(a);
        if (cur->arity == 1) cur = nullptr;
    }
    return p;
}

std::string join_args(int argc, char **argv) {  // std::env::args().join(" "), terminal.rs:300,412
    std::string s;
    for (int i = 0; i < argc; i++) s += (i ? " " : "") + std::string(argv[i]);
    return s;
}

int device_from_env() {
    const char *d = getenv("SCRUBBY_GPU_DEVICE");
    return d ? atoi(d) : 0;
}

}  // namespace

int main(int argc, char **argv) {
    try {
        int i = 1;
        while (i < argc && (!strcmp(argv[i], "-l") || !strcmp(argv[i], "--log-file"))) i += 2;  // terminal.rs:29-30
        if (i >= argc) usage_error("a subcommand is required");
        std::string cmd = argv[i++];
        if (cmd == "--version" || cmd == "-V") {
            printf("scrubby %s\n", CRATE_VERSION);
            return 0;
        }
        if (cmd == "--help" || cmd == "-h" || cmd == "help") {
            fputs(HELP_TOP, stdout);
            return 0;
        }
        for (int k = i; k < argc; k++) {
            if (!strcmp(argv[k], "-A") || !strcmp(argv[k], "--aligner-args") || !strcmp(argv[k], "-C") ||
                !strcmp(argv[k], "--classifier-args")) {
                k++;  // allow_hyphen_values: the next argument is a value, whatever it looks like
                continue;
            }
            if ((!strcmp(argv[k], "--help") || !strcmp(argv[k], "-h")) && help_of(cmd)) {
                fputs(help_of(cmd), stdout);
                return 0;
            }
        }
        if (cmd == "classifier") {
            // note: in the reference both --reads and --json claim -j (terminal.rs:235,259); as in clap's
            // release behaviour the later definition (json) wins for the short form, so use --reads.
            std::vector<Flag> flags = {{'i', "input", 2}, {'o', "output", 2}, {'k', "report", 1}, {'j', "reads", 1},
                                       {'c', "classifier", 1}, {'T', "taxa", 2}, {'D', "taxa-direct", 2}, {'j', "json", 1},
                                       {'w', "workdir", 1}, {'r', "read-ids", 1}, {'e', "extract", 0}};
            Parsed p = parse(argc, argv, i, flags);
            if (!p.one("report")) usage_error("the following required arguments were not provided: --report <REPORT>");
            if (!p.one("reads")) usage_error("the following required arguments were not provided: --reads <READS>");
            if (!p.one("classifier")) usage_error("the following required arguments were not provided: --classifier <CLASSIFIER>");
            auto cls = parse_classifier(*p.one("classifier"));
            if (!cls) usage_error("invalid value '" + *p.one("classifier") + "' for '--classifier' [possible values: kraken2, metabuli]");
            Scrubby s;
            s.input = p.many("input");
            s.output = p.many("output");
            s.json = p.one("json");
            s.workdir = p.one("workdir");
            s.read_ids = p.one("read-ids");
            s.extract = p.has("extract");
            s.device = device_from_env();
            s.config.command = join_args(argc, argv);
            s.config.classifier = *cls;
            s.config.reads = p.one("reads");
            s.config.report = p.one("report");
            s.config.taxa = p.many("taxa");
            s.config.taxa_direct = p.many("taxa-direct");
            build_classifier(s).clean();
        } else if (cmd == "alignment") {
            std::vector<Flag> flags = {{'i', "input", 2}, {'o', "output", 2}, {'a', "alignment", 1}, {'f', "format", 1},
                                       {'l', "min-len", 1}, {'c', "min-cov", 1}, {'q', "min-mapq", 1}, {'j', "json", 1},
                                       {'w', "workdir", 1}, {'r', "read-ids", 1}, {'e', "extract", 0}};
            Parsed p = parse(argc, argv, i, flags);
            if (!p.one("alignment")) usage_error("the following required arguments were not provided: --alignment <ALIGNMENT>");
            Scrubby s;
            s.input = p.many("input");
            s.output = p.many("output");
            s.json = p.one("json");
            s.workdir = p.one("workdir");
            s.read_ids = p.one("read-ids");
            s.extract = p.has("extract");
            s.device = device_from_env();
            s.config.command = join_args(argc, argv);
            s.config.alignment = p.one("alignment");
            if (auto f = p.one("format")) {
                auto fmt = parse_alignment_format(*f);
                if (!fmt) usage_error("invalid value '" + *f + "' for '--format' [possible values: sam, bam, cram, paf, txt, gaf]");
                s.config.alignment_format = fmt;
            }
            try {
                s.config.min_query_length = std::stoull(p.one("min-len").value_or("0"));
                s.config.min_query_coverage = std::stod(p.one("min-cov").value_or("0"));
                unsigned long q = std::stoul(p.one("min-mapq").value_or("0"));
                if (q > 255) throw std::out_of_range("u8");
                s.config.min_mapq = (uint8_t)q;
            } catch (const std::exception &) {
                usage_error("invalid numeric value for --min-len / --min-cov / --min-mapq");
            }
            build_alignment(s).clean();
        } else if (cmd == "diff") {
            std::vector<Flag> flags = {{'i', "input", 2}, {'o', "output", 2}, {'j', "json", 1}, {'r', "read-ids", 1}};
            Parsed p = parse(argc, argv, i, flags);
            ReadDifference d = ReadDifference::build(p.many("input"), p.many("output"), p.one("json"), p.one("read-ids"));
            d.device = device_from_env();
            d.compute();
        } else if (cmd == "reads") {  // terminal.rs:57-157, 176-202
            std::vector<Flag> flags = {{'i', "input", 2}, {'o', "output", 2}, {'I', "index", 1}, {'a', "aligner", 1},
                                       {'p', "preset", 1}, {'c', "classifier", 1}, {'T', "taxa", 2}, {'D', "taxa-direct", 2},
                                       {'A', "aligner-args", 1, true}, {'C', "classifier-args", 1, true}, {'t', "threads", 1},
                                       {'j', "json", 1}, {'w', "workdir", 1}, {'r', "read-ids", 1}, {'e', "extract", 0}};
            Parsed p = parse(argc, argv, i, flags);
            if (!p.one("index")) usage_error("the following required arguments were not provided: --index <INDEX>");
            Scrubby s;
            s.input = p.many("input");
            s.output = p.many("output");
            s.json = p.one("json");
            s.workdir = p.one("workdir");
            s.read_ids = p.one("read-ids");
            s.extract = p.has("extract");
            s.device = device_from_env();
            s.config.command = join_args(argc, argv);
            s.config.index = p.one("index");
            if (auto a = p.one("aligner")) {
                s.config.aligner = parse_aligner(*a);
                if (!s.config.aligner) usage_error("invalid value '" + *a + "' for '--aligner' [possible values: bowtie2, minimap2, minigraph, strobealign]");
            }
            if (auto c = p.one("classifier")) {
                s.config.classifier = parse_classifier(*c);
                if (!s.config.classifier) usage_error("invalid value '" + *c + "' for '--classifier' [possible values: kraken2, metabuli]");
            }
            if (auto pr = p.one("preset")) {
                s.config.preset = parse_preset(*pr);
                if (!s.config.preset) usage_error("invalid value '" + *pr + "' for '--preset'");
            }
            s.config.taxa = p.many("taxa");
            s.config.taxa_direct = p.many("taxa-direct");
            s.config.aligner_args = p.one("aligner-args");
            s.config.classifier_args = p.one("classifier-args");
            try {
                s.threads = (unsigned)std::stoul(p.one("threads").value_or("4"));
            } catch (const std::exception &) {
                usage_error("invalid value for '--threads'");
            }
            build(s).clean();
        } else {
            usage_error("unrecognized subcommand '" + cmd + "'");
        }
    } catch (const ScrubbyError &e) {
        fprintf(stderr, "Error: %s\n", e.what());  // anyhow-style: non-zero exit (main.rs:8,42)
        return 1;
    }
    return 0;
}
