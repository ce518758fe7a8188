#include "image_io.hpp"
#include "image_pool.hpp"
#include <cstdio>
#include <exception>
int main(int argc, char** argv) {
    int ok = 0, err = 0;
    for (int i = 1; i < argc; ++i) {
        int e[2] = {0, 0}, ch = 0;
        try {
            uint8_t* px = dlimg::load_image(argv[i], e, &ch);
            // touch every byte
            size_t n = (size_t)e[0] * e[1] * ch; unsigned s = 0;
            for (size_t k = 0; k < n; ++k) s += px[k];
            dlimg::image_free(px);
            ++ok; (void)s;
        } catch (std::exception const&) { ++err; }
    }
    std::printf("ok %d err %d\n", ok, err);
    return 0;
}
