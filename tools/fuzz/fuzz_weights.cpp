#include "weights.hpp"
#include <cstdio>
#include <exception>
int main(int argc, char** argv) {
    int ok = 0, err = 0;
    for (int i = 1; i < argc; ++i) {
        try {
            dlimg::WeightFile wf = dlimg::WeightFile::load(argv[i]);
            (void)wf.get("a.weight");
            ++ok;
        } catch (std::exception const&) { ++err; }
    }
    std::printf("ok %d err %d\n", ok, err);
}
