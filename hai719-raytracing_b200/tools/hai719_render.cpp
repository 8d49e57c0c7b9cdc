// hai719_render — the headless stand-in for "press r" in the reference's GLUT shell
// (main.cpp:321-326 -> ray_trace_from_camera). Builds one of the reference's scenes with the host
// API, renders it on the GPU and writes the same P3 rendu.ppm.
//   hai719_render [--scene N | --scene-file FILE] [--w W --h H] [--spp S] [--seed K] [--assets DIR] [--out FILE] [--p6 1 | --png 1] [--exr FILE] [--device D | --gpus N]
// --gpus N: render on devices 0..N-1 of this box at once (tiles round-robin, one framebuffer on device 0 written over NVLink)
//                 [--preview PASSES [--orbit PIXELS]]
// --preview: progressive refinement instead of one render (host/Preview.h): PASSES passes of --spp samples each, every
// intermediate frame written as <out>.<pass>.ppm (binary); with --orbit the left mouse button is "dragged" PIXELS to
// the right after every second pass, which restarts the frame like any camera move in the reference's window would.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include "Camera.h"
#include "Constants.h"
#include "Preview.h"
#include "Renderer.h"
#include "Scene.h"
#include "SceneFile.h"

int main(int argc, char **argv) {
    int scene_id = DEFAULT_SELECTED_SCENE, w = 850, h = 480, spp = DEFAULT_NSAMPLES, device = 0;
    unsigned int seed = 0;
    std::string assets, out = "./rendu.ppm", scene_file;
    bool p6 = false, png = false;
    std::string exr;   // float render: also write the linear image as OpenEXR
    int preview = 0, orbit = 0, gpus = 1;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i];
        const char *v = argv[i + 1];
        if (k == "--scene") scene_id = std::atoi(v);
        else if (k == "--w") w = std::atoi(v);
        else if (k == "--h") h = std::atoi(v);
        else if (k == "--spp") spp = std::atoi(v);
        else if (k == "--seed") seed = (unsigned int)std::strtoul(v, nullptr, 10);
        else if (k == "--assets") assets = v;
        else if (k == "--out") out = v;
        else if (k == "--device") device = std::atoi(v);
        else if (k == "--gpus") gpus = std::atoi(v);
        else if (k == "--scene-file") scene_file = v;
        else if (k == "--p6") p6 = std::atoi(v) != 0;
        else if (k == "--png") png = std::atoi(v) != 0;
        else if (k == "--exr") exr = v;
        else if (k == "--preview") preview = std::atoi(v);
        else if (k == "--orbit") orbit = std::atoi(v);
        else { std::cerr << "unknown option " << k << std::endl; return 2; }
    }
    Scene scene;
    scene.asset_root = assets;
    seed_scene_random(seed);
    if (!scene_file.empty()) {
        std::string err;
        if (!hai719::load_scene_file(scene, scene_file, &err)) { std::cerr << err << std::endl; return 2; }
    } else if (!scene.setup_by_id(scene_id, float(w) / float(h))) { std::cerr << "unknown scene " << scene_id << std::endl; return 2; }
    Camera camera;
    camera.resize(w, h);
    camera.move(0., 0., -3.1);   // main.cpp:418
    hai719::RenderOptions opt;
    opt.seed = seed;
    opt.device = device;
    if (gpus > 1) for (int d = 0; d < gpus; ++d) opt.devices.push_back(d);
    opt.ppm_path = out;
    opt.exr_path = exr;
    std::vector<Vec3> image;
    try {
        if (preview > 0) {
            hai719::DeviceScene dev(scene, device);
            opt.verbose = false;
            hai719::Preview pv(dev, camera, w, h, opt);
            for (int pass = 1; pass <= preview; ++pass) {
                RtStats st;
                const unsigned int n = pv.pass((unsigned int)spp, &st);
                const std::string name = out + "." + std::to_string(pass) + ".ppm";
                if (!pv.save(name)) { std::cerr << "Could not open file: " << name << std::endl; return 1; }
                std::cout << "pass " << pass << ": " << n << " samples per pixel, " << st.kernel_ms << " ms on the GPU -> " << name << std::endl;
                if (orbit != 0 && pass % 2 == 0 && pass < preview) {
                    pv.mouse(hai719::Preview::Left, hai719::Preview::Down, w / 2, h / 2);
                    pv.motion(w / 2 + orbit, h / 2);
                    pv.mouse(hai719::Preview::Left, hai719::Preview::Up, w / 2 + orbit, h / 2);
                }
            }
        } else if (p6 || png) {   // output stage on the GPU: 8-bit quantise on the device, binary PPM or PNG
            opt.format = png ? hai719::RenderOptions::PNG : hai719::RenderOptions::P6;
            hai719::DeviceScene dev(scene, device);
            std::vector<unsigned char> rgb8;
            hai719::ray_trace_from_camera_rgb8(dev, camera, w, h, (unsigned int)spp, rgb8, opt);
        } else
        hai719::ray_trace_from_camera(scene, camera, w, h, (unsigned int)spp, image, opt);
    } catch (const std::exception &e) {
        std::cerr << "render failed: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
