// dropin_client.cpp -- an application written against the REFERENCE's unmodified public C++ header
// (/root/reference/src/include/dlimgedit/dlimgedit.hpp, README.md:17-33 usage) and linked against this
// repo's libdlimgedit.so.  It proves the drop-in boundary: same header, same calls, new engine.
// Built by __graft_entry__.build() where /root/reference exists; run by tests/test_gpu_dropin.py.
//
// usage: dropin_client <model_dir> <raw_rgba_file> <width> <height> <out_prefix>
#include <dlimgedit/dlimgedit.hpp>
#include <dlimg_b200.hpp>  // this repo: C++ wrappers of the additive batch table, next to the reference façade

#include <cstdio>
#include <fstream>
#include <iostream>
#include <vector>

using namespace dlimg;

static void dump(Image const& img, std::string const& path) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<char const*>(img.pixels()), (std::streamsize)img.size());
}

int main(int argc, char** argv) {
    if (argc != 6) return 2;
    std::string const model_dir = argv[1], prefix = argv[5];
    int const w = std::atoi(argv[3]), h = std::atoi(argv[4]);
    std::vector<uint8_t> pixels((size_t)w * h * 4);
    std::ifstream(argv[2], std::ios::binary).read(reinterpret_cast<char*>(pixels.data()), (std::streamsize)pixels.size());

    try {
        std::cout << "cpu supported: " << Environment::is_supported(Backend::cpu) << "\n";
        std::cout << "gpu supported: " << Environment::is_supported(Backend::gpu) << "\n";
        Options opts;
        opts.backend = Backend::gpu;
        opts.model_directory = model_dir.c_str();
        Environment env(opts);

        ImageView view(pixels.data(), Extent{w, h}, Channels::rgba);
        Segmentation seg = Segmentation::process(view, env);
        std::cout << "extent: " << seg.extent().width << "x" << seg.extent().height << "\n";

        Image mask = seg.compute_mask(Point{w / 3, h / 2});
        dump(mask, prefix + "_point.raw");
        Image box = seg.compute_mask(Region(Point{w / 8, h / 8}, Extent{w / 2, h / 2}));
        dump(box, prefix + "_region.raw");
        auto masks = seg.compute_masks(Point{w / 3, h / 2});
        for (int i = 0; i < 3; ++i) {
            dump(masks[i].image, prefix + "_multi" + std::to_string(i) + ".raw");
            std::cout << "accuracy " << i << ": " << masks[i].accuracy << "\n";
        }
        Image::save(mask, (prefix + "_point.png").c_str());
        Image reloaded = Image::load((prefix + "_point.png").c_str());
        std::cout << "png roundtrip: " << (reloaded.size() == mask.size() &&
                                           std::equal(mask.pixels(), mask.pixels() + mask.size(), reloaded.pixels()))
                  << "\n";

        // the additive extension through include/dlimg_b200.hpp: a batch of two images in one encoder pass, prompts of both
        // images in one decoder pass; results must equal the one-at-a-time reference calls above
        {
            std::vector<uint8_t> flipped(pixels.rbegin(), pixels.rend());
            ImageView views[2] = {view, ImageView(flipped.data(), Extent{w, h}, Channels::rgba)};
            std::vector<Segmentation> segs = b200::process_batch(env, views, 2);
            Segmentation const* owners[3] = {&segs[0], &segs[1], &segs[0]};
            b200::Prompt prompts[3] = {Point{w / 3, h / 2}, Point{w / 2, h / 2}, Region(Point{w / 8, h / 8}, Extent{w / 2, h / 2})};
            std::vector<Image> out;
            for (int i = 0; i < 3; ++i) out.emplace_back(Extent{w, h}, Channels::mask);
            uint8_t* ptrs[3] = {out[0].pixels(), out[1].pixels(), out[2].pixels()};
            float ious[3] = {0, 0, 0};
            b200::compute_masks_batch(env, owners, prompts, 3, false, ptrs, ious);
            bool const same = std::equal(mask.pixels(), mask.pixels() + mask.size(), out[0].pixels()) &&
                              std::equal(box.pixels(), box.pixels() + box.size(), out[2].pixels());
            Image second = segs[1].compute_mask(Point{w / 2, h / 2});
            bool const same2 = std::equal(second.pixels(), second.pixels() + second.size(), out[1].pixels());
            std::cout << "batch extension: " << (same && same2) << " launches " << b200::stats(env).kernel_launches << "\n";
        }

        // error path: exceptions carry last_error() text (dlimgedit.impl.hpp:7-11)
        try {
            Options bad;
            bad.backend = Backend::gpu;
            bad.model_directory = "/nonexistent/models";
            Environment e2(bad);
            std::cout << "error: not raised\n";
        } catch (Exception const& e) {
            std::cout << "error: " << e.what() << "\n";
        }
    } catch (std::exception const& e) {
        std::cerr << "FAILED: " << e.what() << "\n";
        return 1;
    }
    std::cout << "done\n";
    return 0;
}
