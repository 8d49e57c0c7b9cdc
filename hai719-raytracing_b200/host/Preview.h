// Interactive preview — the part of the GLUT shell (main.cpp:266-437) that surrounds the renderer,
// without the window: the reference's mouse handlers drive the reference's Camera, and instead of one blocking
// ray_trace_from_camera() per 'r' key press the frame is refined pass by pass on the GPU (rt_accum_*), each pass
// adding samples to the ones already there. A viewer (any toolkit, a socket, a file sequence) calls
//     mouse()/motion()/resize()   from its input callbacks      (same arguments as GLUT's, main.cpp:344-388)
//     pass()                      from its idle callback         (traces `pass_spp` more samples per pixel)
//     frame_rgb8()                from its display callback      (current mean, 8-bit, row 0 = top)
// Moving the camera restarts the accumulation. After passes of s1, s2, ... samples with the camera at rest the frame
// equals ray_trace_from_camera() at nsamples = s1 + s2 + ... bit for bit (tests/test_gpu_parity.py).
#ifndef HAI719_HOST_PREVIEW_H
#define HAI719_HOST_PREVIEW_H
#include <vector>
#include "Renderer.h"

namespace hai719 {

class Preview {
public:
    enum Button { Left = 0, Middle = 1, Right = 2 };        // GLUT_LEFT_BUTTON, GLUT_MIDDLE_BUTTON, GLUT_RIGHT_BUTTON
    enum State { Down = 0, Up = 1 };                        // GLUT_DOWN, GLUT_UP

    // `camera` is the caller's (the reference keeps one global Camera, main.cpp:43); it must outlive the preview.
    // The preview holds its own reference on the scene's device arrays: destroying or re-uploading the DeviceScene
    // while the preview lives is safe (it goes on refining the scene it was created on).
    Preview(const DeviceScene &scene, Camera &camera, int w, int h, const RenderOptions &opt = RenderOptions());
    ~Preview();
    Preview(const Preview &) = delete;
    Preview &operator=(const Preview &) = delete;

    void mouse(int button, int state, int x, int y);        // main.cpp:344-372
    void motion(int x, int y);                              // main.cpp:374-388 (screen size = this preview's w, h)
    void resize(int w, int h);                              // main.cpp:391-394 (Camera::resize; new accumulator)
    void invalidate();                                      // camera or scene changed behind the preview's back

    unsigned int pass(unsigned int pass_spp = 1, RtStats *stats = nullptr);   // returns samples per pixel so far
    unsigned int samples() const;
    int width() const { return w_; }
    int height() const { return h_; }

    const std::vector<unsigned char> &frame_rgb8();         // (int)(255.f*min(1.f,c)) of the current mean, main.cpp:258
    void frame(std::vector<Vec3> &image);                   // gamma-corrected floats, as ray_trace_from_camera()'s `image`
    bool save(const std::string &filename, RenderOptions::Format format = RenderOptions::P6);

private:
    void rebuild();
    RtScene *scene_;                                        // retained (rt_scene_retain): the device copy outlives its DeviceScene if need be
    Camera &camera_;
    RenderOptions opt_;
    int w_, h_;
    RtAccum *accum_ = nullptr;
    bool dirty_ = true;                                     // camera changed since the last pass
    bool rotate_ = false, move_ = false, zoom_ = false;     // mouseRotatePressed / mouseMovePressed / mouseZoomPressed
    int last_x_ = 0, last_y_ = 0, last_zoom_ = 0;
    std::vector<unsigned char> rgb8_;
    unsigned int rgb8_samples_ = 0;                         // sample count rgb8_ was read at (0 = stale)
};

}  // namespace hai719
#endif
