// Sketch.h -- host-side mirror of the reference's public Sketch surface (reference src/mash/Sketch.h:28-224),
// re-hosted on the C ABI of libmashgpu (include/mashgpu.h).  Same member names, argument meaning, stderr messages
// and exit(1) behaviour, so that callers written against the reference's Sketch.h keep compiling; the hashing,
// bottom-s selection, comparison and screening all happen on the GPU.  Windowed sketches (COMMAND_FIND), Bloom /
// min-copies / target-coverage read filters are not provided (dead by default / "next").
#ifndef MASHHOST_SKETCH_H
#define MASHHOST_SKETCH_H

#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mashgpu.h"

namespace mash {

static const char *const alphabetNucleotide = "ACGT";                   // reference Sketch.h:25
static const char *const alphabetProtein = "ACDEFGHIKLMNPQRSTVWY";      // reference Sketch.h:26
static const char *const suffixSketch = ".msh";

typedef uint32_t hash32_t;
typedef uint64_t hash64_t;
union hash_u { hash32_t hash32; hash64_t hash64; };                     // reference hash.h:15-19

// reference HashList.h:13-37 (one vector: 32-bit hashes are kept widened, as the C ABI carries them)
class HashList {
public:
    HashList() : use64(true) {}
    explicit HashList(bool use64new) : use64(use64new) {}
    hash_u at(int index) const { hash_u h; h.hash64 = 0; if (use64) h.hash64 = hashes[index]; else h.hash32 = (hash32_t)hashes[index]; return h; }
    void clear() { hashes.clear(); }
    void resize(int size) { hashes.resize(size); }
    void set32(int index, uint32_t value) { hashes[index] = value; }
    void set64(int index, uint64_t value) { hashes[index] = value; }
    void setUse64(bool use64New) { use64 = use64New; }
    int size() const { return (int)hashes.size(); }
    void push_back32(hash32_t hash) { hashes.push_back(hash); }
    void push_back64(hash64_t hash) { hashes.push_back(hash); }
    bool get64() const { return use64; }
    const uint64_t *data() const { return hashes.data(); }
private:
    bool use64;
    std::vector<uint64_t> hashes;
};

class Sketch {
public:
    typedef uint64_t hash_t;

    struct Parameters {   // reference Sketch.h:34-109
        Parameters()
            : parallelism(1), kmerSize(0), alphabetSize(0), preserveCase(false), use64(false), seed(0), error(0), warning(0),
              minHashesPerWindow(0), windowSize(0), windowed(false), concatenated(false), noncanonical(false), reads(false),
              memoryBound(0), minCov(1), targetCov(0), genomeSize(0), counts(false)
        { memset(alphabet, 0, 256); }
        int parallelism;
        int kmerSize;
        bool alphabet[256];
        uint32_t alphabetSize;
        bool preserveCase;
        bool use64;
        uint32_t seed;
        double error;
        double warning;
        uint64_t minHashesPerWindow;
        uint64_t windowSize;
        bool windowed;
        bool concatenated;
        bool noncanonical;
        bool reads;
        uint64_t memoryBound;
        uint32_t minCov;
        double targetCov;
        uint64_t genomeSize;
        bool counts;
    };

    struct Reference {    // reference Sketch.h:131-139
        std::string name;
        std::string comment;
        uint64_t length = 0;
        HashList hashesSorted;
        std::vector<uint32_t> counts;
        bool countsSorted = false;
    };

    void getAlphabetAsString(std::string &alphabet) const;
    uint32_t getAlphabetSize() const { return parameters.alphabetSize; }
    bool getConcatenated() const { return parameters.concatenated; }
    float getError() const { return parameters.error; }
    uint32_t getHashSeed() const { return parameters.seed; }
    float getMinHashesPerWindow() const { return parameters.minHashesPerWindow; }
    int getMinKmerSize(uint64_t reference) const;
    bool getPreserveCase() const { return parameters.preserveCase; }
    double getRandomKmerChance(uint64_t reference) const;
    const Reference &getReference(uint64_t index) const { return references.at(index); }
    uint64_t getReferenceCount() const { return references.size(); }
    void getReferenceHistogram(uint64_t index, std::map<uint32_t, uint64_t> &histogram) const;
    uint64_t getReferenceIndex(std::string id) const;
    int getKmerSize() const { return parameters.kmerSize; }
    double getKmerSpace() const { return kmerSpace; }
    bool getUse64() const { return parameters.use64; }
    uint64_t getWindowSize() const { return parameters.windowSize; }
    bool getNoncanonical() const { return parameters.noncanonical; }
    bool hasHashCounts() const { return references.size() > 0 && references.at(0).counts.size() > 0; }
    int initFromFiles(const std::vector<std::string> &files, const Parameters &parametersNew, int verbosity = 0, bool enforceParameters = false, bool contain = false);
    void initFromReads(const std::vector<std::string> &files, const Parameters &parametersNew);
    uint64_t initParametersFromCapnp(const char *file);
    void setReferenceName(int i, const std::string name) { references[i].name = name; }
    void setReferenceComment(int i, const std::string comment) { references[i].comment = comment; }
    int writeToCapnp(const char *file) const;

    // fills a mashgpu_sketch_set view of this sketch (rows of `stride` hashes); the vectors own the storage
    void toSketchSet(mashgpu_sketch_set &set, std::vector<uint64_t> &hashes, std::vector<uint32_t> &n, std::vector<uint64_t> &lengths) const;

    const Parameters &getParameters() const { return parameters; }
    // assembling a Sketch from parts (mash paste, tests): not in the reference's surface
    void setParameters(const Parameters &parametersNew) { parameters = parametersNew; createIndex(); }
    void addReference(const Reference &reference) { references.push_back(reference); createIndex(); }

private:
    struct Batch;
    void flushBatch(Batch &batch);
    void loadCapnp(const char *file);
    void createIndex();

    std::vector<Reference> references;
    std::unordered_map<std::string, int> referenceIndecesById;
    Parameters parameters;
    double kmerSpace = 0;
};

void setAlphabetFromString(Sketch::Parameters &parameters, const char *characters);
bool hasSuffix(std::string const &whole, std::string const &suffix);
mashgpu_ctx *gpuContext();     // one engine context per process (device from MASH_GPU_DEVICE, default 0)
void gpuContextBegin();        // start creating it on a helper thread (gpuContext() then waits for that)
void fillGpuParams(mashgpu_sketch_params &p, const Sketch::Parameters &parameters);

}  // namespace mash

#endif
