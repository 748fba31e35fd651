// Options.h -- the reference's option surface for the k-mer spectrum path, same flag names and defaults:
//   GeneralOptions                 src/Options.h:325-657
//   KmerBaseOptions                src/Kmer.h:132-165
//   KmerSpectrumOptions            src/KmerSpectrum.h:90-256
//   ReadSelectorOptions            src/ReadSelector.h:70-204
//   FilterReadsBaseOptions         apps/FilterReads.h:101-150
//   FilterKnownOdditiesOptions     src/FilterKnownOddities.h:73-134
//   MPIOptions                     src/MPIBuffer.h:66-99            (accepted; the exchange is NCCL here)
//   DuplicateFragmentFilterOptions src/DuplicateFragmentFilter.h:60-113 (accepted; only dedup-mode 0 is supported)
// The reference composes one static singleton per group and parses with boost::program_options, which accepts
// unambiguous prefixes (its tests pass --thread and --out); this parser keeps both properties.
// Positional arguments: kmer-size, then input-file... (apps/FilterReads.cpp:68-69).
// README spellings --min-kmer-depth / --max-kmer-depth (README.md:125) are aliases of --min-depth /
// --max-kmer-output-depth.
#ifndef KMERNATOR_HOST_OPTIONS_H
#define KMERNATOR_HOST_OPTIONS_H

#include <cstdlib>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "Log.h"

class OptionsBaseInterface {
public:
    typedef std::vector<std::string> FileListType;
};

namespace kmn_host {

struct OptionSpec {
    std::string name, help, value, def;
    bool isList = false, isSet = false, supported = true;
    std::vector<std::string> values;
};

class OptionRegistry {
public:
    static OptionRegistry &get() { static OptionRegistry r; return r; }
    void add(const std::string &name, const std::string &def, const std::string &help, bool isList = false, bool supported = true)
    {
        if (index.count(name)) return;
        OptionSpec s; s.name = name; s.def = def; s.value = def; s.help = help; s.isList = isList; s.supported = supported;
        index[name] = specs.size();
        specs.push_back(s);
    }
    void alias(const std::string &from, const std::string &to) { aliases[from] = to; }
    OptionSpec &spec(const std::string &name)
    {
        std::map<std::string, size_t>::iterator it = index.find(name);
        if (it == index.end()) LOG_THROW("unknown option: " << name);
        return specs[it->second];
    }
    // exact name, alias, or unambiguous prefix (boost::program_options allow_guessing)
    std::string resolve(const std::string &given)
    {
        if (aliases.count(given)) return aliases[given];
        if (index.count(given)) return given;
        std::string found;
        for (size_t i = 0; i < specs.size(); ++i)
            if (specs[i].name.compare(0, given.size(), given) == 0) {
                if (!found.empty()) LOG_THROW("option '--" << given << "' is ambiguous (" << found << ", " << specs[i].name << ")");
                found = specs[i].name;
            }
        if (found.empty()) LOG_THROW("unrecognised option '--" << given << "'");
        return found;
    }
    void set(const std::string &name, const std::string &v)
    {
        OptionSpec &s = spec(name);
        if (s.isList) s.values.push_back(v); else s.value = v;
        s.isSet = true;
    }
    std::string usage() const
    {
        std::ostringstream ss;
        for (size_t i = 0; i < specs.size(); ++i) {
            ss << "  --" << specs[i].name;
            if (!specs[i].def.empty()) ss << " arg (=" << specs[i].def << ")";
            ss << "\t" << specs[i].help << (specs[i].supported ? "" : "  [accepted; not implemented by kmernator_b200]") << "\n";
        }
        return ss.str();
    }
    std::vector<OptionSpec> specs;
private:
    std::map<std::string, size_t> index;
    std::map<std::string, std::string> aliases;
};

inline long asLong(const std::string &n) { return std::strtol(OptionRegistry::get().spec(n).value.c_str(), NULL, 10); }
inline double asDouble(const std::string &n) { return std::strtod(OptionRegistry::get().spec(n).value.c_str(), NULL); }
inline std::string asString(const std::string &n) { return OptionRegistry::get().spec(n).value; }

}  // namespace kmn_host

// ---- option groups: accessor names follow the reference's getters -------------------------------------------
class _GeneralOptions : public OptionsBaseInterface {
public:
    FileListType inputFiles;
    std::vector<std::string> inputFilePrefixes;
    static void setOptions()
    {
        kmn_host::OptionRegistry &r = kmn_host::OptionRegistry::get();
        r.add("help", "", "produce help message");
        r.add("verbose", "1", "level of verbosity (0+)");
        r.add("debug", "0", "level of debug verbosity (0+)");
        r.add("log-file", "", "If set all INFO and DEBUG messages will be logged here", false, false);
        r.add("gathered-logs", "1", "(MPI only) gather logs to the master", false, false);
        r.add("threads", "0", "maximum number of host threads (the count pass runs on the GPU)");
        r.add("batch-size", "100000", "reads per batch handed to the GPU");
        r.add("input-file", "", "input file(s)", true);
        r.add("output-file", "", "output file pattern");
        r.add("mmap-input", "1", "mmap input files", false, false);
        r.add("fastq-base-quality", "33", "ASCII value for quality 0 in the input (auto-detected)");
        r.add("fastq-output-base-quality", "33", "ASCII value for quality 0 in the output");
        r.add("ignore-quality", "0", "ignore the quality score, to save memory or if they are untrusted");
        r.add("min-quality-score", "3", "minimum quality score over entire kmer");
        r.add("format-output", "0", "0: fastq, 1: fasta, 2: fastq unmasked, 3: fasta unmasked");
        r.add("keep-read-comment", "1", "preserve the comment part of the read header");
        r.add("build-output-in-memory", "0", "build output in memory before writing", false, false);
        r.add("temp-dir", "/tmp", "temporary directory", false, false);
        r.add("keep-temp-dir", "", "keep temporary directory", false, false);
    }
    FileListType &getInputFiles() { return inputFiles; }
    std::string getOutputFile() { return kmn_host::asString("output-file"); }
    int getVerbose() { return (int)kmn_host::asLong("verbose"); }
    int getDebug() { return (int)kmn_host::asLong("debug"); }
    unsigned int getBatchSize() { return (unsigned int)kmn_host::asLong("batch-size"); }
    int getFastqBaseQuality() { return (int)kmn_host::asLong("fastq-base-quality"); }
    int getOutputFastqBaseQuality() { return (int)kmn_host::asLong("fastq-output-base-quality"); }
    bool getIgnoreQual() { return kmn_host::asLong("ignore-quality") != 0; }
    unsigned int getMinQuality() { return (unsigned int)kmn_host::asLong("min-quality-score"); }
    int getFormatOutput() { return (int)kmn_host::asLong("format-output"); }
    bool getKeepReadComment() { return kmn_host::asLong("keep-read-comment") != 0; }
    // basename of the input file up to its last '.'                      src/Options.h:531-551
    std::string &getInputFileSubstring(unsigned int fileIdx)
    {
        if (inputFilePrefixes.empty()) {
            for (FileListType::iterator it = inputFiles.begin(); it != inputFiles.end(); ++it) {
                size_t start = it->find_last_of('/');
                start = (start == std::string::npos) ? 0 : start + 1;
                size_t end = it->find_last_of('.');
                if (end == std::string::npos) end = it->length() - 1;
                inputFilePrefixes.push_back(it->substr(start, end - start));
            }
        }
        return inputFilePrefixes[fileIdx];
    }
};
class Options {
public:
    static _GeneralOptions &getOptions() { static _GeneralOptions o; return o; }
};

class _KmerBaseOptions {
public:
    static void setOptions()
    {
        kmn_host::OptionRegistry &r = kmn_host::OptionRegistry::get();
        r.add("kmer-size", "23", "kmer size.  A size of 0 will skip k-mer calculations");   // src/KmerSpectrum.h:99-101
        r.add("kmers-per-bucket", "32", "number of kmers to target per hash-bucket (sizing hint only)");
    }
    unsigned int getKmerSize() { return (unsigned int)kmn_host::asLong("kmer-size"); }
};
class KmerBaseOptions { public: static _KmerBaseOptions &getOptions() { static _KmerBaseOptions o; return o; } };

class _KmerSpectrumOptions {
public:
    static void setOptions()
    {
        kmn_host::OptionRegistry &r = kmn_host::OptionRegistry::get();
        r.add("min-kmer-quality", "0.10", "minimum quality-adjusted kmer probability (0-1)");
        r.add("min-depth", "2", "minimum depth for a solid kmer");
        r.alias("min-kmer-depth", "min-depth");
        r.add("estimated-depth", "20", "sizing hint", false, true);
        r.add("estimated-error-rate", "0.35", "sizing hint", false, true);
        r.add("save-kmer-mmap", "0", "If set, creates a memory map of the kmer spectrum for later use");
        r.add("load-kmer-mmap", "", "Instead of generating kmer spectrum, load an existing one named by this option");
        r.add("build-partitions", "0", "build the spectrum in this many hash partitions (one pass suffices in HBM)");
        r.add("kmer-subsample", "1", "subsample kmers", false, false);
        r.add("variant-sigmas", "-1", "purge variants", false, false);
        r.add("min-variant-kmer-depth", "512", "purge variants", false, false);
        r.add("variant-edit-disance", "2", "purge variants", false, false);
        r.add("periodic-singleton-purge", "0", "purge singletons periodically", false, false);
        r.add("gc-heat-map", "0", "GC heat map", false, false);
    }
    double getMinKmerQuality() { return kmn_host::asDouble("min-kmer-quality"); }
    unsigned int getMinDepth() { return (unsigned int)kmn_host::asLong("min-depth"); }
    unsigned int getBuildPartitions() { return (unsigned int)kmn_host::asLong("build-partitions"); }
    bool getSaveKmerMmap() { return kmn_host::asLong("save-kmer-mmap") != 0; }
    std::string getLoadKmerMmap() { return kmn_host::asString("load-kmer-mmap"); }
    double getEstimatedDepth() { return kmn_host::asDouble("estimated-depth"); }
    double getEstimatedErrorRate() { return kmn_host::asDouble("estimated-error-rate"); }
};
class KmerSpectrumOptions { public: static _KmerSpectrumOptions &getOptions() { static _KmerSpectrumOptions o; return o; } };

class _ReadSelectorOptions {
public:
    float minReadLengthOverride;
    int minPassingOverride;
    _ReadSelectorOptions() : minReadLengthOverride(-1.f), minPassingOverride(-1) {}
    static void setOptions()
    {
        kmn_host::OptionRegistry &r = kmn_host::OptionRegistry::get();
        r.add("separate-outputs", "1", "split the output by input file and depth");
        r.add("max-kmer-output-depth", "-1", "maximum number of times a kmer will be output among the selected reads");
        r.alias("max-kmer-depth", "max-kmer-output-depth");
        r.add("use-logscale-above-max", "0", "pick reads above max depth on a log scale");
        r.add("normalization-method", "RANDOM", "RANDOM or OPTIMAL (OPTIMAL is serial-only in the reference; not implemented)");
        r.add("partition-by-depth", "-1", "partition output by depth", false, false);
        r.add("min-passing-in-pair", "1", "1 or 2 reads in a pair must pass filters");
        r.add("min-read-length", "0.40", "minimum (trimmed) read length; <= 1.0 is a fraction of the read length");
        r.add("remainder-trim", "-1", "trim remainder", false, false);
        r.add("kmer-scoring-type", "MAX", "SUM, MEDIAN, AVG, MIN or MAX");
        r.add("bimodal-sigmas", "-1", "bimodal read detection", false, false);
    }
    bool getSeparateOutputs() { return kmn_host::asLong("separate-outputs") != 0; }
    int getMaxKmerDepth() { return (int)kmn_host::asLong("max-kmer-output-depth"); }
    bool getUseLogscaleAboveMax() { return kmn_host::asLong("use-logscale-above-max") != 0; }
    std::string getNormalizationMethod() { return kmn_host::asString("normalization-method"); }
    int getMinPassingInPair() { return (int)kmn_host::asLong("min-passing-in-pair"); }
    bool getBothPairs() { return getMinPassingInPair() == 2; }
    float getMinReadLength() { return (float)kmn_host::asDouble("min-read-length"); }
    std::string getKmerScoringType() { return kmn_host::asString("kmer-scoring-type"); }
};
class ReadSelectorOptions { public: static _ReadSelectorOptions &getOptions() { static _ReadSelectorOptions o; return o; } };

class _FilterKnownOdditiesOptions {
public:
    static void setOptions()
    {
        kmn_host::OptionRegistry &r = kmn_host::OptionRegistry::get();
        r.add("skip-artifact-filter", "0", "skip homo-polymer, primer-dimer and duplicated fragment pair filtering");
        r.add("artifact-match-length", "24", "kmer match length to known artifact sequences (a multiple of 4, at most 28)");
        r.add("artifact-edit-distance", "2", "edit distance to apply to artifact-match-length matches to known artifacts");
        r.add("build-artifact-edits-in-filter", "2", "0 - edits are searched for at run time, 1 - built into the filter, 2 - built while the filter is small");
        r.add("mask-simple-repeats", "0", "mask simple repeats", false, true);
        r.add("phix-output", "0", "separate PhiX reads", false, false);
        r.add("filter-output", "0", "separate artifact reads", false, false);
        r.add("artifact-reference-file", "", "additional artifact reference file(s)", true, false);
    }
    bool getSkipArtifactFilter() { return kmn_host::asLong("skip-artifact-filter") != 0; }
};
class FilterKnownOdditiesOptions { public: static _FilterKnownOdditiesOptions &getOptions() { static _FilterKnownOdditiesOptions o; return o; } };

class _FilterReadsBaseOptions {
public:
    static void setOptions()
    {
        kmn_host::OptionRegistry &r = kmn_host::OptionRegistry::get();
        r.add("histogram-file", "", "if set, the kmer histogram is written to this file");
        r.add("size-history-file", "", "if set, the kmer spectrum size history (for EstimateSize.R) is written here");
        r.add("subtract-file", "", "file(s) whose abundant kmers (>= min-depth) are subtracted from the spectrum", true);
        r.add("reference-file", "", "reference file(s) whose kmers are subtracted from the spectrum", true);
        // MPIOptions / DuplicateFragmentFilterOptions: accepted so reference command lines parse
        r.add("mpi-buffer-size", "33554432", "accepted; the exchange is an NCCL all-to-all here");
        r.add("mpi-min-transmit-size", "2048", "accepted; unused");
        r.add("dedup-mode", "0", "0 = no fragment de-duplication (the only supported mode)");
        r.add("dedup-single", "0", "dedup", false, false);
        r.add("dedup-consensus", "1", "dedup", false, false);
        r.add("dedup-edit-distance", "0", "dedup", false, false);
        r.add("dedup-start-offset", "0", "dedup", false, false);
        r.add("dedup-length", "24", "dedup", false, false);
    }
    std::string getHistogramFile() { return kmn_host::asString("histogram-file"); }
    std::string getSizeHistoryFile() { return kmn_host::asString("size-history-file"); }
    OptionsBaseInterface::FileListType getSubtractFiles() { return kmn_host::OptionRegistry::get().spec("subtract-file").values; }
    OptionsBaseInterface::FileListType getReferenceFiles() { return kmn_host::OptionRegistry::get().spec("reference-file").values; }
};
class FilterReadsBaseOptions { public: static _FilterReadsBaseOptions &getOptions() { static _FilterReadsBaseOptions o; return o; } };

// FilterReadsOptions::parseOpts (apps/FilterReads.cpp:57-81, apps/FilterReads.h:113-150)
class FilterReadsOptions {
public:
    static void registerAll()
    {
        _GeneralOptions::setOptions();
        _KmerBaseOptions::setOptions();
        _KmerSpectrumOptions::setOptions();
        _ReadSelectorOptions::setOptions();
        _FilterKnownOdditiesOptions::setOptions();
        _FilterReadsBaseOptions::setOptions();
    }
    // resetDefaults: an app's own defaults (e.g. MeraculousCounter: min-quality-score 2, min-kmer-quality 0,
    // apps/MeraculousCounter.cpp:66-80), applied before the command line
    static void setDefault(const std::string &name, const std::string &v)
    {
        kmn_host::OptionSpec &s = kmn_host::OptionRegistry::get().spec(name);
        s.def = v; s.value = v;
    }
    static bool parseOpts(int argc, char *argv[], void (*resetDefaults)() = NULL, bool kmerSizePositional = true)
    {
        registerAll();
        if (resetDefaults) resetDefaults();
        kmn_host::OptionRegistry &r = kmn_host::OptionRegistry::get();
        try {
            std::vector<std::string> positional;
            for (int i = 1; i < argc; ++i) {
                std::string a = argv[i];
                if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
                    std::string name = a.substr(2), val;
                    bool hasVal = false;
                    size_t eq = name.find('=');
                    if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); hasVal = true; }
                    name = r.resolve(name);
                    if (name == "help") { std::cerr << usage(argv[0]); return false; }
                    if (!hasVal) {
                        if (i + 1 >= argc) LOG_THROW("option '--" << name << "' requires an argument");
                        val = argv[++i];
                    }
                    r.set(name, val);
                } else if (a == "-h") { std::cerr << usage(argv[0]); return false; }
                else positional.push_back(a);
            }
            size_t p = 0;
            if (kmerSizePositional && !r.spec("kmer-size").isSet && p < positional.size()) r.set("kmer-size", positional[p++]);
            for (; p < positional.size(); ++p) r.set("input-file", positional[p]);
            Options::getOptions().inputFiles = r.spec("input-file").values;
            Log::verboseLevel() = Options::getOptions().getVerbose();
            Log::debugLevel() = Options::getOptions().getDebug();
            // refuse what this build does not implement instead of silently ignoring it
            for (size_t i = 0; i < r.specs.size(); ++i)
                if (r.specs[i].isSet && !r.specs[i].supported && r.specs[i].value != r.specs[i].def)
                    LOG_THROW("--" << r.specs[i].name << " is part of the reference's option surface but is not implemented by kmernator_b200");
            if (kmn_host::asLong("dedup-mode") != 0) LOG_THROW("--dedup-mode > 0 is not implemented by kmernator_b200");
            if (Options::getOptions().getInputFiles().empty()) LOG_THROW("Please specify at least one input file");
            int ob = Options::getOptions().getOutputFastqBaseQuality();
            if (ob != 33 && ob != 64) LOG_THROW("--fastq-output-base-quality must be 33 or 64");
        } catch (std::exception &e) {
            std::cerr << usage(argv[0]) << "\n" << e.what() << std::endl;
            return false;
        }
        return true;
    }
    static std::string usage(const char *argv0)
    {
        return std::string("Usage: ") + argv0 + " [options] kmer-size input-file [input-file ...]\n" + kmn_host::OptionRegistry::get().usage();
    }
};

#endif
