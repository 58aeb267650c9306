// mash.cpp -- thin command-line front end over the host shim: `mash sketch | dist | triangle | screen | info`.
// Mirrors the reference's command drivers (CommandSketch.cpp:36-169, CommandDistance.cpp:44-304,
// CommandTriangle.cpp:45-214, CommandScreen.cpp:54-461, CommandInfo.cpp:222-299): same option identifiers and
// defaults (Command.cpp:165-200), same stdout formats (iostream default precision), same ordering contracts.
// It exists so that the reference's `make test` flow can be replayed against the GPU engine; help text, `paste`,
// `bounds`, `taxscreen`, `within`, `find`, winner-take-all and translated screening are out of scope.
#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "Sketch.h"
#include <unistd.h>

#include "fastout.hpp"
#include "fastx.hpp"

using namespace std;
using namespace mash;

namespace {

// reference Command::Option (Command.h:28-69, Command.cpp:48-156): numbers are parsed with stof into a float
struct Option {
    enum Type { Boolean, Number, Integer, Size, File, String } type;
    string identifier, argument;
    float number = 0;
    float argMin = 0, argMax = 0;
    bool active = false;
    Option() : type(Boolean) {}
    Option(Type t, const string &id, const string &def, float lo = 0, float hi = 0) : type(t), identifier(id), argMin(lo), argMax(hi) { set(def); }
    void set(string a)
    {
        argument = a;
        if (type == Number || type == Integer) {
            if (a.empty()) { number = 0; return; }
            bool failed = false;
            try {
                number = stof(a);
                if (argMin != argMax && (number < argMin || number > argMax)) failed = true;
                else if (type == Integer && uint64_t(number) != number) failed = true;
            } catch (const exception &) { failed = true; }
            if (failed) {
                cerr << "ERROR: Argument to -" << identifier << " must be a" << (type == Integer ? "n integer" : " number");
                if (argMin != argMax) cerr << " between " << argMin << " and " << argMax;
                cerr << " (" << a << " given)" << endl;
                exit(1);
            }
        } else if (type == Size) {
            if (a.empty()) { number = 0; return; }
            char suffix = a[a.size() - 1];
            uint64_t factor = 1;
            if (suffix < '0' || suffix > '9') {
                switch (suffix) {
                    case 'k': case 'K': factor = 1000; break;
                    case 'm': case 'M': factor = 1000000; break;
                    case 'g': case 'G': factor = 1000000000; break;
                    case 't': case 'T': factor = 1000000000000; break;
                    default:
                        cerr << "ERROR: Unrecognized unit (\"" << suffix << "\") in argument to -" << identifier << ". If specified, unit must be one of [kKmMgGtT]." << endl;
                        exit(1);
                }
                a.resize(a.size() - 1);
            }
            bool fail = false;
            try { number = stof(a); } catch (const exception &) { fail = true; }
            if (number <= 0 || (uint64_t)number != number) fail = true;
            if (fail) {
                cerr << "ERROR: Argument to -" << identifier << " must be a whole number, optionally followed by one of [kKmMgGtT]." << endl;
                exit(1);
            }
            number *= factor;
        }
    }
};

struct Command {
    string name;
    map<string, Option> options;          // by name
    map<string, string> byIdentifier;
    vector<string> arguments;

    void add(const string &n, const Option &o) { options[n] = o; byIdentifier[o.identifier] = n; }
    const Option &opt(const string &n) const { return options.at(n); }
    bool has(const string &n) const { return options.count(n) != 0; }

    void useSketchOptions()   // reference Command.cpp:354-379 with the defaults of :165-200
    {
        add("threads", Option(Option::Integer, "p", "1"));
        add("kmer", Option(Option::Integer, "k", "21", 1, 32));
        add("noncanonical", Option(Option::Boolean, "n", ""));
        add("protein", Option(Option::Boolean, "a", ""));
        add("alphabet", Option(Option::String, "z", ""));
        add("case", Option(Option::Boolean, "Z", ""));
        add("sketchSize", Option(Option::Integer, "s", "1000"));
        add("individual", Option(Option::Boolean, "i", ""));
        add("seed", Option(Option::Integer, "S", "42", 0, 0xFFFFFFFF));
        add("warning", Option(Option::Number, "w", "0.01", 0, 1));
        add("reads", Option(Option::Boolean, "r", ""));
        add("memory", Option(Option::Size, "b", ""));
        add("minCov", Option(Option::Integer, "m", "1"));
        add("targetCov", Option(Option::Number, "c", ""));
        add("genome", Option(Option::Size, "g", ""));
    }

    int parse(int argc, const char **argv)   // reference Command::run(argc, argv), Command.cpp:311-347
    {
        for (int i = 0; i < argc; i++) {
            if (argv[i][0] == '-' && argv[i][1] != 0) {
                if (byIdentifier.count(argv[i] + 1) == 0) {
                    cerr << "ERROR: Unrecognized option: " << argv[i] << endl;
                    return 1;
                }
                Option &option = options.at(byIdentifier.at(argv[i] + 1));
                option.active = true;
                if (option.type != Option::Boolean) {
                    i++;
                    if (i == argc) {
                        cerr << "ERROR: -" << option.identifier << " requires an argument" << endl;
                        return 1;
                    }
                    option.set(argv[i]);
                }
            } else {
                arguments.push_back(argv[i]);
            }
        }
        return 0;
    }
};

// reference sketchParameterSetup.cpp:15-105
int sketchParameterSetup(Sketch::Parameters &parameters, const Command &command)
{
    parameters.kmerSize = command.opt("kmer").number;
    parameters.minHashesPerWindow = command.opt("sketchSize").number;
    parameters.concatenated = !command.opt("individual").active;
    parameters.noncanonical = command.opt("noncanonical").active;
    parameters.seed = command.opt("seed").number;
    parameters.reads = command.opt("reads").active;
    parameters.minCov = command.opt("minCov").number;
    parameters.targetCov = command.opt("targetCov").number;
    parameters.parallelism = command.opt("threads").number;
    parameters.preserveCase = command.opt("case").active;
    if (command.has("warning")) parameters.warning = command.opt("warning").number;
    if (command.opt("memory").active) {
        parameters.reads = true;
        parameters.memoryBound = command.opt("memory").number;
        if (command.opt("minCov").active) {
            cerr << "ERROR: The option " << command.opt("minCov").identifier << " cannot be used with " << command.opt("memory").identifier << "." << endl;
            return 1;
        }
    }
    if (command.opt("minCov").active || command.opt("targetCov").active) parameters.reads = true;
    if (command.opt("genome").active) {
        parameters.reads = true;
        parameters.genomeSize = command.opt("genome").number;
    }
    if (parameters.reads) parameters.counts = true;
    if (parameters.reads && command.opt("threads").active)
        cerr << "WARNING: The option " << command.opt("threads").identifier << " will be ignored with " << command.opt("reads").identifier << "." << endl;
    if (parameters.reads && !parameters.concatenated) {
        cerr << "ERROR: The option " << command.opt("individual").identifier << " cannot be used with " << command.opt("reads").identifier << "." << endl;
        return 1;
    }
    if (command.opt("protein").active) {
        parameters.noncanonical = true;
        setAlphabetFromString(parameters, alphabetProtein);
        if (!command.opt("kmer").active) parameters.kmerSize = 9;
        setAlphabetFromString(parameters, alphabetProtein);   // use64 depends on the final k
    } else if (command.opt("alphabet").active) {
        parameters.noncanonical = true;
        setAlphabetFromString(parameters, command.opt("alphabet").argument.c_str());
    } else {
        setAlphabetFromString(parameters, alphabetNucleotide);
    }
    return 0;
}

void warnKmerSize(const Sketch::Parameters &parameters, uint64_t lengthMax, const string &lengthMaxName, double randomChance, int kMin, int warningCount)
{
    cerr << "\nWARNING: For the k-mer size used (" << parameters.kmerSize << "), the random match probability (" << randomChance
         << ") is above the specified warning threshold (" << parameters.warning << ") for the sequence \"" << lengthMaxName << "\" of size " << lengthMax;
    if (warningCount > 1) cerr << " (and " << (warningCount - 1) << " others)";
    cerr << ". Distances to " << (warningCount == 1 ? "this sequence" : "these sequences")
         << " may be underestimated as a result. To meet the threshold of " << parameters.warning << ", a k-mer size of at least " << kMin
         << " is required. See: -k, -w." << endl << endl;
}

void splitFile(const string &file, vector<string> &lines)   // reference Command.cpp splitFile
{
    ifstream in(file);
    if (!in) { cerr << "ERROR: could not open " << file << " for reading." << endl; exit(1); }
    string line;
    while (getline(in, line)) lines.push_back(line);
}

struct KmerWarning {
    uint64_t lengthMax = 0; double randomChance = 0; int kMin = 0; string lengthMaxName; int warningCount = 0;
    void scan(const Sketch &sketch, const Sketch::Parameters &parameters)
    {
        double lengthThreshold = (parameters.warning * sketch.getKmerSpace()) / (1. - parameters.warning);
        for (uint64_t i = 0; i < sketch.getReferenceCount(); i++) {
            uint64_t length = sketch.getReference(i).length;
            if (length > lengthThreshold) {
                if (warningCount == 0 || length > lengthMax) {
                    lengthMax = length; lengthMaxName = sketch.getReference(i).name;
                    randomChance = sketch.getRandomKmerChance(i); kMin = sketch.getMinKmerSize(i);
                }
                warningCount++;
            }
        }
    }
};

void gpuFail()
{
    cerr << "ERROR: " << mashgpu_last_error(gpuContext()) << endl;
    exit(1);
}

// ---------------------------------------------------------------------------------------------------- sketch
int runSketch(int argc, const char **argv)
{
    Command c;
    c.add("help", Option(Option::Boolean, "h", ""));
    c.add("list", Option(Option::Boolean, "l", ""));
    c.add("prefix", Option(Option::File, "o", ""));
    c.add("id", Option(Option::File, "I", ""));
    c.add("comment", Option(Option::File, "C", ""));
    c.add("counts", Option(Option::Boolean, "M", ""));
    c.useSketchOptions();
    if (c.parse(argc, argv)) return 1;
    if (c.arguments.size() == 0 || c.opt("help").active) { cerr << "usage: mash sketch [options] <input> [<input>] ..." << endl; return 0; }
    Sketch::Parameters parameters;
    parameters.counts = c.opt("counts").active;
    if (sketchParameterSetup(parameters, c)) return 1;
    vector<string> files;
    for (auto &a : c.arguments) { if (c.opt("list").active) splitFile(a, files); else files.push_back(a); }
    if (c.opt("id").active || c.opt("comment").active)
        if (files.size() > 1 && !parameters.reads) cerr << "WARNING: -I and -C will only apply to first sketch" << endl;
    Sketch sketch;
    const bool trace = getenv("MASHGPU_TRACE") != 0;
    const auto tStart = std::chrono::steady_clock::now();
    auto elapsed = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count(); };
    if (parameters.reads) sketch.initFromReads(files, parameters);
    else sketch.initFromFiles(files, parameters, trace ? 0 : 1);
    if (trace) cerr << "[mash] parse + sketch of " << files.size() << " files: " << elapsed() << " ms" << endl;
    if (c.opt("id").active) sketch.setReferenceName(0, c.opt("id").argument);
    if (c.opt("comment").active) sketch.setReferenceComment(0, c.opt("comment").argument);
    KmerWarning w;
    w.scan(sketch, parameters);
    string prefix = c.opt("prefix").argument.length() > 0 ? c.opt("prefix").argument : (c.arguments[0] == "-" ? string("stdin") : c.arguments[0]);
    if (!hasSuffix(prefix, suffixSketch)) prefix += suffixSketch;
    cerr << "Writing to " << prefix << "..." << endl;
    sketch.writeToCapnp(prefix.c_str());
    if (trace) cerr << "[mash] .msh written at " << elapsed() << " ms" << endl;
    if (w.warningCount > 0 && !parameters.reads) warnKmerSize(parameters, w.lengthMax, w.lengthMaxName, w.randomChance, w.kMin, w.warningCount);
    return 0;
}

// ---------------------------------------------------------------------------------------------------- dist
struct PairBlock {
    vector<uint32_t> numer, denom; vector<double> distance, pValue; vector<uint8_t> pass;
    void resize(size_t n) { numer.resize(n); denom.resize(n); distance.resize(n); pValue.resize(n); pass.resize(n); }
};

int runDist(int argc, const char **argv)
{
    Command c;
    c.add("help", Option(Option::Boolean, "h", ""));
    c.add("list", Option(Option::Boolean, "l", ""));
    c.add("table", Option(Option::Boolean, "t", ""));
    c.add("pvalue", Option(Option::Number, "v", "1.0", 0., 1.));
    c.add("distance", Option(Option::Number, "d", "1.0", 0., 1.));
    c.add("comment", Option(Option::Boolean, "C", ""));
    c.useSketchOptions();
    if (c.parse(argc, argv)) return 1;
    if (c.arguments.size() < 2 || c.opt("help").active) { cerr << "usage: mash dist [options] <reference> <query> [<query>] ..." << endl; return 0; }
    bool table = c.opt("table").active, comment = c.opt("comment").active;
    double pValueMax = c.opt("pvalue").number, distanceMax = c.opt("distance").number;
    Sketch::Parameters parameters;
    if (sketchParameterSetup(parameters, c)) return 1;
    Sketch sketchRef;
    const string &fileReference = c.arguments[0];
    bool isSketch = hasSuffix(fileReference, suffixSketch);
    if (isSketch) {
        for (const char *o : {"kmer", "noncanonical", "protein", "alphabet"})
            if (c.opt(o).active) {
                cerr << "ERROR: The option -" << c.opt(o).identifier << " cannot be used when a sketch is provided; it is inherited from the sketch." << endl;
                return 1;
            }
    } else {
        cerr << "Sketching " << fileReference << " (provide sketch file made with \"mash sketch\" to skip)...";
    }
    vector<string> refArgVector(1, fileReference);
    sketchRef.initFromFiles(refArgVector, parameters);
    KmerWarning w;
    if (isSketch) {
        if (c.opt("sketchSize").active && parameters.reads && parameters.minHashesPerWindow != sketchRef.getMinHashesPerWindow()) {
            cerr << "ERROR: The sketch size must match the reference when using a bloom filter (leave this option out to inherit from the reference sketch)." << endl;
            return 1;
        }
        parameters.minHashesPerWindow = sketchRef.getMinHashesPerWindow();
        parameters.kmerSize = sketchRef.getKmerSize();
        parameters.noncanonical = sketchRef.getNoncanonical();
        parameters.preserveCase = sketchRef.getPreserveCase();
        parameters.seed = sketchRef.getHashSeed();
        string alphabet;
        sketchRef.getAlphabetAsString(alphabet);
        setAlphabetFromString(parameters, alphabet.c_str());
    } else {
        w.scan(sketchRef, parameters);
        cerr << "done.\n";
    }
    if (table) {
        cout << "#query";
        for (uint64_t i = 0; i < sketchRef.getReferenceCount(); i++) cout << '\t' << sketchRef.getReference(i).name;
        cout << endl;
    }
    vector<string> queryFiles;
    for (size_t i = 1; i < c.arguments.size(); i++) { if (c.opt("list").active) splitFile(c.arguments[i], queryFiles); else queryFiles.push_back(c.arguments[i]); }
    Sketch sketchQuery;
    sketchQuery.initFromFiles(queryFiles, parameters, 0, true);

    const uint64_t nRef = sketchRef.getReferenceCount(), nQry = sketchQuery.getReferenceCount();
    if (nRef && nQry) {
        uint64_t sketchSize = sketchQuery.getMinHashesPerWindow() < sketchRef.getMinHashesPerWindow() ? sketchQuery.getMinHashesPerWindow() : sketchRef.getMinHashesPerWindow();
        mashgpu_dist_params dp = {sketchSize, sketchRef.getKmerSize(), sketchRef.getKmerSpace(), distanceMax, pValueMax};
        mashgpu_sketch_set setRef, setQry;
        vector<uint64_t> hR, lR, hQ, lQ; vector<uint32_t> nR, nQ;
        sketchRef.toSketchSet(setRef, hR, nR, lR);
        sketchQuery.toSketchSet(setQry, hQ, nQ, lQ);
        mashgpu_dist_job *job = 0;
        if (mashgpu_dist_open(gpuContext(), &setRef, &setQry, &dp, &job) != MASHGPU_OK) gpuFail();
        uint64_t rows = std::max<uint64_t>(1, (1ull << 24) / nRef);
        PairBlock out;
        out.resize(std::min(rows, nQry) * nRef);
        // with -d / -v filters only the passing pairs are needed (writeOutput prints nothing else): ask the engine for the
        // compacted pass-list, already in query-major order, instead of the dense grid
        const bool filtered = !table && (distanceMax < 1.0 || pValueMax < 1.0);
        const uint64_t listCapacity = 1ull << 22;
        vector<uint64_t> lIdx; vector<uint32_t> lNumer, lDenom; vector<double> lDist, lP;
        if (filtered) { lIdx.resize(listCapacity); lNumer.resize(listCapacity); lDenom.resize(listCapacity); lDist.resize(listCapacity); lP.resize(listCapacity); }
        for (uint64_t q = 0; q < nQry; q += rows) {
            uint64_t r = std::min(rows, nQry - q);
            if (filtered) {
                uint64_t nPass = 0;
                if (mashgpu_dist_run_list(job, q, r, listCapacity, lIdx.data(), lNumer.data(), lDenom.data(), lDist.data(), lP.data(), &nPass) != MASHGPU_OK) gpuFail();
                if (nPass <= listCapacity) {
                    mashhost::writeRows(nPass, parameters.parallelism, 1, [&](uint64_t e, mashhost::OutBuf &o) {
                        uint64_t i = lIdx[e] / nRef, j = lIdx[e] % nRef;
                        o.str(sketchRef.getReference(j).name);
                        if (comment) { o.ch(':'); o.str(sketchRef.getReference(j).comment); }
                        o.ch('\t'); o.str(sketchQuery.getReference(q + i).name);
                        if (comment) { o.ch(':'); o.str(sketchQuery.getReference(q + i).comment); }
                        o.ch('\t'); o.dbl(lDist[e]); o.ch('\t'); o.dbl(lP[e]); o.ch('\t'); o.u64(lNumer[e]); o.ch('/'); o.u64(lDenom[e]); o.ch('\n');
                    });
                    continue;
                }
                // more passing pairs than the list holds: dense path for this block
            }
            if (mashgpu_dist_run(job, q, r, out.numer.data(), out.denom.data(), out.distance.data(), out.pValue.data(), out.pass.data()) != MASHGPU_OK) gpuFail();
            mashhost::writeRows(r, parameters.parallelism, nRef, [&](uint64_t i, mashhost::OutBuf &o) {   // == writeOutput, reference CommandDistance.cpp:247-304
                for (uint64_t j = 0; j < nRef; j++) {
                    size_t k = i * nRef + j;
                    if (table && j == 0) o.str(sketchQuery.getReference(q + i).name);
                    if (table) {
                        o.ch('\t');
                        if (out.pass[k]) o.dbl(out.distance[k]);
                    } else if (out.pass[k]) {
                        o.str(sketchRef.getReference(j).name);
                        if (comment) { o.ch(':'); o.str(sketchRef.getReference(j).comment); }
                        o.ch('\t'); o.str(sketchQuery.getReference(q + i).name);
                        if (comment) { o.ch(':'); o.str(sketchQuery.getReference(q + i).comment); }
                        o.ch('\t'); o.dbl(out.distance[k]); o.ch('\t'); o.dbl(out.pValue[k]); o.ch('\t'); o.u64(out.numer[k]); o.ch('/'); o.u64(out.denom[k]); o.ch('\n');
                    }
                }
                if (table) o.ch('\n');
            });
        }
        mashgpu_dist_close(job);
    }
    if (w.warningCount > 0 && !parameters.reads) warnKmerSize(parameters, w.lengthMax, w.lengthMaxName, w.randomChance, w.kMin, w.warningCount);
    return 0;
}

// ---------------------------------------------------------------------------------------------------- triangle
int runTriangle(int argc, const char **argv)
{
    Command c;
    c.add("help", Option(Option::Boolean, "h", ""));
    c.add("list", Option(Option::Boolean, "l", ""));
    c.add("comment", Option(Option::Boolean, "C", ""));
    c.add("edge", Option(Option::Boolean, "E", ""));
    c.add("pvalue", Option(Option::Number, "v", "1.0", 0., 1.));
    c.add("distance", Option(Option::Number, "d", "1.0", 0., 1.));
    c.useSketchOptions();
    if (c.parse(argc, argv)) return 1;
    if (c.arguments.size() < 1 || c.opt("help").active) { cerr << "usage: mash triangle [options] <seq1> [<seq2>] ..." << endl; return 0; }
    bool comment = c.opt("comment").active, edge = c.opt("edge").active;
    double pValueMax = c.opt("pvalue").number, distanceMax = c.opt("distance").number, pValuePeak = 0;
    if (c.opt("pvalue").active || c.opt("distance").active) edge = true;
    Sketch::Parameters parameters;
    if (sketchParameterSetup(parameters, c)) return 1;
    if (c.arguments.size() == 1 && !c.opt("list").active) parameters.concatenated = false;
    vector<string> queryFiles;
    for (auto &a : c.arguments) { if (c.opt("list").active) splitFile(a, queryFiles); else queryFiles.push_back(a); }
    Sketch sketch;
    sketch.initFromFiles(queryFiles, parameters);
    KmerWarning w;
    w.scan(sketch, parameters);
    const uint64_t n = sketch.getReferenceCount();
    if (!edge) {
        cout << '\t' << n << endl;
        cout << (comment ? sketch.getReference(0).comment : sketch.getReference(0).name) << endl;
    }
    if (n > 1) {
        // row i vs 0..i-1 (reference CommandTriangle.cpp:200-214): reference = sketch i, "query" = sketch j < i.
        // compareSketches(ref = row sketch, qry = column sketch); shared counts and distance are symmetric, the
        // p-value is symmetric in the two lengths, so the all-vs-all job's (q = i, r = j) entry is the same value.
        mashgpu_dist_params dp = {(uint64_t)sketch.getMinHashesPerWindow(), sketch.getKmerSize(), sketch.getKmerSpace(), distanceMax, pValueMax};
        mashgpu_sketch_set set;
        vector<uint64_t> h, l; vector<uint32_t> nn;
        sketch.toSketchSet(set, h, nn, l);
        mashgpu_dist_job *job = 0;
        if (mashgpu_dist_open(gpuContext(), &set, 0, &dp, &job) != MASHGPU_OK) gpuFail();
        if (mashgpu_dist_set_triangle(job, 1) != MASHGPU_OK) gpuFail();     // only j < i is read below
        uint64_t rows = std::max<uint64_t>(1, (1ull << 24) / n);
        PairBlock out;
        out.resize(std::min(rows, n) * n);
        for (uint64_t q = 1; q < n; q += rows) {
            uint64_t r = std::min(rows, n - q);
            if (mashgpu_dist_run(job, q, r, out.numer.data(), out.denom.data(), out.distance.data(), out.pValue.data(), out.pass.data()) != MASHGPU_OK) gpuFail();
            mashhost::writeRows(r, parameters.parallelism, n, [&](uint64_t i, mashhost::OutBuf &o) {   // == writeOutput, reference CommandTriangle.cpp:159-198
                const Sketch::Reference &ref = sketch.getReference(q + i);
                if (!edge) o.str(comment ? ref.comment : ref.name);
                for (uint64_t j = 0; j < q + i; j++) {
                    size_t k = i * n + j;
                    if (edge) {
                        if (out.pass[k]) {
                            const Sketch::Reference &qry = sketch.getReference(j);
                            o.str(comment ? ref.comment : ref.name); o.ch('\t'); o.str(comment ? qry.comment : qry.name); o.ch('\t');
                            o.dbl(out.distance[k]); o.ch('\t'); o.dbl(out.pValue[k]); o.ch('\t'); o.u64(out.numer[k]); o.ch('/'); o.u64(out.denom[k]); o.ch('\n');
                        }
                    } else {
                        o.ch('\t'); o.dbl(out.distance[k]);
                    }
                }
                if (!edge) o.ch('\n');
            });
            for (uint64_t i = 0; i < r; i++)
                for (uint64_t j = 0; j < q + i; j++)
                    if (out.pValue[i * n + j] > pValuePeak) pValuePeak = out.pValue[i * n + j];
        }
        mashgpu_dist_close(job);
    }
    if (!edge) cerr << "Max p-value: " << pValuePeak << endl;
    if (w.warningCount > 0 && !parameters.reads) warnKmerSize(parameters, w.lengthMax, w.lengthMaxName, w.randomChance, w.kMin, w.warningCount);
    return 0;
}

// ---------------------------------------------------------------------------------------------------- screen
int runScreen(int argc, const char **argv)
{
    Command c;
    c.add("help", Option(Option::Boolean, "h", ""));
    c.add("threads", Option(Option::Integer, "p", "1"));
    c.add("winning!", Option(Option::Boolean, "w", ""));
    c.add("identity", Option(Option::Number, "i", "0", -1., 1.));
    c.add("pvalue", Option(Option::Number, "v", "1.0", 0., 1.));
    if (c.parse(argc, argv)) return 1;
    if (c.arguments.size() < 2 || c.opt("help").active) { cerr << "usage: mash screen [options] <queries>.msh <mixture> [<mixture>] ..." << endl; return 0; }
    if (!hasSuffix(c.arguments[0], suffixSketch)) {
        cerr << "ERROR: " << c.arguments[0] << " does not look like a sketch (.msh)" << endl;
        exit(1);
    }
    double pValueMax = c.opt("pvalue").number, identityMin = c.opt("identity").number;
    Sketch sketch;
    Sketch::Parameters parameters;
    vector<string> refArgVector(1, c.arguments[0]);
    sketch.initFromFiles(refArgVector, parameters);
    parameters = sketch.getParameters();
    string alphabet;
    sketch.getAlphabetAsString(alphabet);
    if (alphabet == alphabetProtein) { cerr << "ERROR: translated (amino acid) screening is not available in the GPU engine." << endl; return 1; }
    cerr << "Loading " << c.arguments[0] << "..." << endl;
    mashgpu_sketch_params p;
    fillGpuParams(p, parameters);
    mashgpu_sketch_set set;
    vector<uint64_t> h, l; vector<uint32_t> nn;
    sketch.toSketchSet(set, h, nn, l);
    mashgpu_screen_job *job = 0;
    if (mashgpu_screen_open(gpuContext(), &p, &set, &job) != MASHGPU_OK) gpuFail();
    if (c.opt("winning!").active && mashgpu_screen_set_winner(job, 1) != MASHGPU_OK) gpuFail();     // -w, CommandScreen.cpp:357-407
    int queryCount = (int)c.arguments.size() - 1;
    cerr << "Streaming from ";
    if (queryCount == 1) cerr << c.arguments[1]; else cerr << queryCount << " inputs";
    cerr << "..." << endl;
    // round robin over the mixture files, '*'-joined chunks (reference CommandScreen.cpp:156-270; chunks of 64 MiB
    // instead of 1 MiB -- any split gives the same counts)
    vector<gzFile> fps;
    std::vector<mashhost::FastxReader *> readers;
    for (size_t f = 1; f < c.arguments.size(); f++) {
        if (c.arguments[f] == "-" && f > 1) { cerr << "ERROR: '-' for stdin must be first query" << endl; exit(1); }
        gzFile fp = mashhost::FastxReader::openPath(c.arguments[f]);
        if (fp == 0) { cerr << "ERROR: could not open " << c.arguments[f] << endl; exit(1); }
        fps.push_back(fp);
        readers.push_back(new mashhost::FastxReader(fp));
    }
    const size_t chunkSize = 64u << 20;
    string input;
    input.reserve(chunkSize + (1 << 20));
    uint64_t count = 0;
    int lstate = 0;
    size_t it = 0;
    const int kmerSize = parameters.kmerSize;
    auto flush = [&]() {
        if (input.empty()) return;
        if (mashgpu_screen_feed(job, input.data(), input.size()) != MASHGPU_OK) gpuFail();
        input.clear();
    };
    while (!readers.empty()) {
        int len = readers[it]->read();
        lstate = len;
        if (len < -1) break;
        if (len == -1) {
            delete readers[it];
            readers.erase(readers.begin() + it);
            if (it >= readers.size()) it = 0;
            continue;
        }
        if (input.length() + (len >= kmerSize ? len + 1 : 0) > chunkSize) flush();
        count++;
        if (len >= kmerSize) {
            input.append(1, '*');
            input.append(readers[it]->seq.data(), readers[it]->seq.size());
        }
        it++;
        if (it >= readers.size()) it = 0;
    }
    flush();
    if (lstate != -1) { cerr << "\nERROR: reading inputs" << endl; exit(1); }
    for (gzFile fp : fps) gzclose(fp);
    if (count == 0) { cerr << "\nERROR: Did not find sequence records in inputs" << endl; exit(1); }
    const uint64_t n = sketch.getReferenceCount();
    vector<uint64_t> shared(n), median(n); vector<double> identity(n), pValue(n);
    uint64_t setSize = 0;
    if (mashgpu_screen_finish(job, shared.data(), median.data(), identity.data(), pValue.data(), &setSize, 0, 0) != MASHGPU_OK) gpuFail();
    mashgpu_screen_close(job);
    cerr << "   Estimated distinct k-mers in mixture: " << setSize << endl;
    if (setSize == 0) cerr << "WARNING: no valid k-mers in input." << endl;
    cerr << "Summing shared..." << endl << "Computing coverage medians..." << endl << "Writing output..." << endl;
    for (uint64_t i = 0; i < n; i++) {                       // reference CommandScreen.cpp:418-455
        if (shared[i] != 0 || identityMin < 0.0) {
            if (identity[i] < identityMin) continue;
            if (pValue[i] > pValueMax) continue;
            cout << identity[i] << '\t' << shared[i] << '/' << sketch.getReference(i).hashesSorted.size() << '\t' << median[i] << '\t' << pValue[i]
                 << '\t' << sketch.getReference(i).name << '\t' << sketch.getReference(i).comment << endl;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------- info
int runInfo(int argc, const char **argv)
{
    Command c;
    c.add("help", Option(Option::Boolean, "h", ""));
    c.add("header", Option(Option::Boolean, "H", ""));
    c.add("tabular", Option(Option::Boolean, "t", ""));
    c.add("counts", Option(Option::Boolean, "c", ""));
    c.add("dump", Option(Option::Boolean, "d", ""));
    if (c.parse(argc, argv)) return 1;
    if (c.arguments.size() == 0 || c.opt("help").active) { cerr << "usage: mash info [options] <sketch>" << endl; return 0; }
    const string &file = c.arguments[0];
    if (!hasSuffix(file, suffixSketch)) { cerr << "ERROR: The file \"" << file << "\" does not look like a sketch." << endl; return 1; }
    Sketch sketch;
    Sketch::Parameters params;
    uint64_t referenceCount = sketch.initParametersFromCapnp(file.c_str());
    params = sketch.getParameters();
    if (c.opt("header").active) {
        string alphabet;
        sketch.getAlphabetAsString(alphabet);
        cout << "Header:" << endl;
        cout << "  Hash function (seed):          " << "MurmurHash3_x64_128" << " (" << sketch.getHashSeed() << ")" << endl;
        cout << "  K-mer size:                    " << sketch.getKmerSize() << " (" << (sketch.getUse64() ? "64" : "32") << "-bit hashes)" << endl;
        cout << "  Alphabet:                      " << alphabet << (sketch.getNoncanonical() ? "" : " (canonical)") << (sketch.getPreserveCase() ? " (case-sensitive)" : "") << endl;
        cout << "  Target min-hashes per sketch:  " << sketch.getMinHashesPerWindow() << endl;
        cout << "  Sketches:                      " << referenceCount << endl;
        return 0;
    }
    vector<string> files(1, file);
    sketch.initFromFiles(files, params);
    if (c.opt("dump").active) {                              // writeJson, reference CommandInfo.cpp:222-299
        string alphabet;
        sketch.getAlphabetAsString(alphabet);
        bool use64 = sketch.getUse64();
        cout << "{" << endl;
        cout << "	\"kmer\" : " << sketch.getKmerSize() << ',' << endl;
        cout << "	\"alphabet\" : \"" << alphabet << "\"," << endl;
        cout << "	\"preserveCase\" : " << (sketch.getPreserveCase() ? "true" : "false") << ',' << endl;
        cout << "	\"canonical\" : " << (sketch.getNoncanonical() ? "false" : "true") << ',' << endl;
        cout << "	\"sketchSize\" : " << sketch.getMinHashesPerWindow() << ',' << endl;
        cout << "	\"hashType\" : \"" << "MurmurHash3_x64_128" << "\"," << endl;
        cout << "	\"hashBits\" : " << (use64 ? 64 : 32) << ',' << endl;
        cout << "	\"hashSeed\" : " << sketch.getHashSeed() << ',' << endl;
        cout << " 	\"sketches\" :" << endl;
        cout << "	[" << endl;
        for (uint64_t i = 0; i < sketch.getReferenceCount(); i++) {
            const Sketch::Reference &ref = sketch.getReference(i);
            cout << "		{" << endl;
            cout << "			\"name\" : \"" << ref.name << "\"," << endl;
            cout << "			\"length\" : " << ref.length << ',' << endl;
            cout << "			\"comment\" : \"" << ref.comment << "\"," << endl;
            cout << "			\"hashes\" :" << endl;
            cout << "			[" << endl;
            for (int j = 0; j < ref.hashesSorted.size(); j++) {
                cout << "				" << (use64 ? ref.hashesSorted.at(j).hash64 : ref.hashesSorted.at(j).hash32);
                if (j < ref.hashesSorted.size() - 1) cout << ',';
                cout << endl;
            }
            cout << "			]" << endl;
            if (ref.countsSorted) {
                cout << "			\"counts\" :" << endl;
                cout << "			[" << endl;
                for (int j = 0; j < (int)ref.counts.size(); j++) {
                    cout << "				" << ref.counts.at(j);
                    if (j < ref.hashesSorted.size() - 1) cout << ',';
                    cout << endl;
                }
                cout << "			]" << endl;
            }
            cout << (i < sketch.getReferenceCount() - 1 ? "		}," : "		}") << endl;
        }
        cout << "	]" << endl;
        cout << "}" << endl;
        return 0;
    }
    // default / -t: one line per sketch: hashes, length, name, comment (reference CommandInfo.cpp:139-183, tabular form)
    if (!c.opt("tabular").active) cout << "#Hashes\tLength\tID\tComment" << endl;
    for (uint64_t i = 0; i < sketch.getReferenceCount(); i++) {
        const Sketch::Reference &ref = sketch.getReference(i);
        cout << ref.hashesSorted.size() << '\t' << ref.length << '\t' << ref.name << '\t' << ref.comment << endl;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------- paste
// reference CommandPaste.cpp:30-88: concatenate sketch files with matching parameters into <out_prefix>.msh
int runPaste(int argc, const char **argv)
{
    Command c;
    c.add("help", Option(Option::Boolean, "h", ""));
    c.add("list", Option(Option::Boolean, "l", ""));
    if (c.parse(argc, argv)) return 1;
    if (c.arguments.size() < 2 || c.opt("help").active) { cerr << "usage: mash paste <out_prefix> <sketch> [<sketch>] ..." << endl; return 0; }
    vector<string> files;
    for (size_t i = 1; i < c.arguments.size(); i++) { if (c.opt("list").active) splitFile(c.arguments[i], files); else files.push_back(c.arguments[i]); }
    for (auto &f : files)
        if (!hasSuffix(f, suffixSketch)) { cerr << "ERROR: The file \"" << f << "\" does not look like a sketch." << endl; return 1; }
    Sketch sketch;
    Sketch::Parameters parameters;
    sketch.initFromFiles(files, parameters);
    string out = c.arguments[0];
    if (!hasSuffix(out, suffixSketch)) out += suffixSketch;
    cerr << "Writing " << out << "..." << endl;
    sketch.writeToCapnp(out.c_str());
    return 0;
}

// test helper: rebuild a .msh from a `mash info -d` JSON dump (the reference's golden files are such dumps)
int runImportJson(int argc, const char **argv)
{
    if (argc < 2) { cerr << "usage: mash import-json <dump.json> <out.msh>" << endl; return 1; }
    ifstream in(argv[0]);
    if (!in) { cerr << "ERROR: could not open " << argv[0] << endl; return 1; }
    string text((istreambuf_iterator<char>(in)), istreambuf_iterator<char>());
    auto scalar = [&](const string &key, size_t from = 0) -> string {
        size_t p = text.find("\"" + key + "\" :", from);
        if (p == string::npos) return "";
        p = text.find(':', p) + 1;
        size_t e = text.find('\n', p);
        string v = text.substr(p, e - p);
        while (!v.empty() && (v.back() == ',' || isspace((unsigned char)v.back()))) v.pop_back();
        size_t b = 0;
        while (b < v.size() && isspace((unsigned char)v[b])) b++;
        v = v.substr(b);
        if (v.size() >= 2 && v.front() == '"' && v.back() == '"') v = v.substr(1, v.size() - 2);
        return v;
    };
    Sketch::Parameters parameters;
    parameters.kmerSize = atoi(scalar("kmer").c_str());
    parameters.preserveCase = scalar("preserveCase") == "true";
    parameters.noncanonical = scalar("canonical") == "false";
    parameters.minHashesPerWindow = strtoull(scalar("sketchSize").c_str(), 0, 10);
    parameters.seed = (uint32_t)strtoul(scalar("hashSeed").c_str(), 0, 10);
    parameters.concatenated = true;
    setAlphabetFromString(parameters, scalar("alphabet").c_str());
    Sketch sketch;
    sketch.setParameters(parameters);
    size_t pos = text.find("\"sketches\"");
    while ((pos = text.find("\"name\" :", pos)) != string::npos) {
        Sketch::Reference ref;
        ref.name = scalar("name", pos);
        ref.length = strtoull(scalar("length", pos).c_str(), 0, 10);
        ref.comment = scalar("comment", pos);
        ref.hashesSorted.setUse64(parameters.use64);
        size_t a = text.find('[', text.find("\"hashes\"", pos)), e = text.find(']', a);
        const char *q = text.c_str() + a + 1, *end = text.c_str() + e;
        while (q < end) {
            while (q < end && !isdigit((unsigned char)*q)) q++;
            if (q >= end) break;
            char *next;
            uint64_t h = strtoull(q, &next, 10);
            ref.hashesSorted.push_back64(h);
            q = next;
        }
        sketch.addReference(ref);
        pos = e;
    }
    sketch.writeToCapnp(argv[1]);
    return 0;
}

}  // namespace

static int runCommand(int argc, const char **argv)
{
    string cmd = argv[1];
    if (cmd == "sketch") return runSketch(argc - 2, argv + 2);
    if (cmd == "dist") return runDist(argc - 2, argv + 2);
    if (cmd == "triangle") return runTriangle(argc - 2, argv + 2);
    if (cmd == "screen") return runScreen(argc - 2, argv + 2);
    if (cmd == "info") return runInfo(argc - 2, argv + 2);
    if (cmd == "paste") return runPaste(argc - 2, argv + 2);
    if (cmd == "import-json") return runImportJson(argc - 2, argv + 2);
    cerr << "ERROR: Unrecognized command: " << cmd << endl;
    return 1;
}

int main(int argc, const char **argv)
{
    if (argc < 2) {
        cerr << "usage: mash <sketch|dist|triangle|screen|info> ...   (B200 engine; see INTEGRATION.md)" << endl;
        return 0;
    }
    const int rc = runCommand(argc, argv);
    // Everything the command wrote is closed or flushed here; leave without the static destructors and the CUDA runtime's
    // teardown (unmapping gigabytes of device and pinned memory one allocation at a time): the driver reclaims it all at once.
    cout.flush(); cerr.flush();
    fflush(stdout); fflush(stderr);
    _exit(rc);
}
