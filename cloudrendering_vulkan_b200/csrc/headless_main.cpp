// headless_main.cpp -- crv_headless: the reference's main.cpp around the PathTracer/ShadowVolume hot path, headless.
//
// Same call order as the interactive app, on the C++ RenderTechnique classes of include/crv/render_technique.h:
//   StartSimulation (main.cpp:927-944): SetCloudProperties, SetFOV, techniques, LoadCloudFile, UpdateCloudData
//   UpdateCloudData (main.cpp:266-335): upload, QueueUpdateCloudData(+Sampler), UpdateDescriptorSets, UpdateShadowVolume
//   UpdateShadowVolume (main.cpp:223-264): SetLightDirection, SetOrigin, SV RecordDrawCommands, rebind on PT
//   RenderLoop / DrawFrame (main.cpp:714-736, 603-712): seed = rand(), RecordDrawCommands, frameCount++
// minus window, swapchain, ImGui and the blit.  One zero-cleared result image, frameCount starting at 0
// (SURVEY.md 3.3).  No Vulkan, no GLFW; every computation runs in libcrv_b200.so on the GPU.
//
//   crv_headless [--xyz file | --fbm N] [--width W --height H] [--spp S] [--nsv N] [--g G] [--absorption A]
//                [--density-scaling D] [--light x,y,z] [--per-frame] [--device i] [--pfm out.pfm]
//                [--gpus N [--same-device]]   the picture's frames split over N GPUs (frame f -> member f mod N), every member
//                                             with its own replica of the grid and the shadow volume, combined by crv_group_*
// prints one JSON line: resolution, spp, ms, Mpath-samples/s, SV build ms, FNV-1a hash of the RGBA32F image.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/crv/grid3d.h"
#include "../../include/crv/render_technique.h"

namespace {

struct Options {
    std::string xyz, pfm, png, resume, checkpoint;
    unsigned fbm = 128;
    int width = 512, height = 512, device = 0;
    unsigned spp = 16, nsv = 500;
    float g = 0.85f, absorption = 0.0f, density_scaling = 50.0f, fov = 90.0f;
    float light[3] = {1.0f, -1.0f, 0.0f};
    bool isotropic = false, per_frame = false, warmup = false, math_driver = false;
    unsigned ppm_frames = 0;          // --ppm N: after the path-traced image, N progressive-photon-mapping frames into a second image
    unsigned gpus = 1;                // --gpus N: the frames of the picture are split over N group members
    bool same_device = false;         // --same-device: all members on --device (functional test of the group on one GPU)
    int skip_level = 1;               // --skip-level 0|1 (crv_pt_set_skip_level)
};

bool parse(int argc, char** argv, Options& o)
{
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "crv_headless: %s needs a value\n", name); std::exit(2); }
            return argv[++i];
        };
        if (a == "--xyz") o.xyz = next("--xyz");
        else if (a == "--fbm") o.fbm = (unsigned)std::atoi(next("--fbm"));
        else if (a == "--width") o.width = std::atoi(next("--width"));
        else if (a == "--height") o.height = std::atoi(next("--height"));
        else if (a == "--spp") o.spp = (unsigned)std::atoi(next("--spp"));
        else if (a == "--nsv") o.nsv = (unsigned)std::atoi(next("--nsv"));
        else if (a == "--g") o.g = (float)std::atof(next("--g"));
        else if (a == "--isotropic") o.isotropic = true;
        else if (a == "--absorption") o.absorption = (float)std::atof(next("--absorption"));
        else if (a == "--density-scaling") o.density_scaling = (float)std::atof(next("--density-scaling"));
        else if (a == "--fov") o.fov = (float)std::atof(next("--fov"));
        else if (a == "--light") { if (std::sscanf(next("--light"), "%f,%f,%f", &o.light[0], &o.light[1], &o.light[2]) != 3) return false; }
        else if (a == "--per-frame") o.per_frame = true;
        else if (a == "--device") o.device = std::atoi(next("--device"));
        else if (a == "--pfm") o.pfm = next("--pfm");
        else if (a == "--png") o.png = next("--png");
        else if (a == "--warmup") o.warmup = true;
        else if (a == "--ppm") o.ppm_frames = (unsigned)std::atoi(next("--ppm"));
        else if (a == "--math") { const std::string v = next("--math"); if (v != "strict" && v != "driver") return false; o.math_driver = v == "driver"; }
        else if (a == "--gpus") o.gpus = (unsigned)std::atoi(next("--gpus"));
        else if (a == "--same-device") o.same_device = true;
        else if (a == "--skip-level") o.skip_level = std::atoi(next("--skip-level"));
        else if (a == "--resume") o.resume = next("--resume");
        else if (a == "--checkpoint") o.checkpoint = next("--checkpoint");
        else return false;
    }
    if (o.gpus > 1 && (o.per_frame || !o.resume.empty() || o.ppm_frames > 0)) {
        std::fprintf(stderr, "crv_headless: --gpus N renders one batch picture (no --per-frame / --resume / --ppm)\n");
        return false;
    }
    return o.width > 0 && o.height > 0 && o.spp > 0 && o.nsv > 0 && o.gpus >= 1 && o.gpus <= 16;
}

uint64_t fnv1a(const void* data, size_t n)
{
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

// the g_* globals of main.cpp that the hot path touches (main.cpp:31-97)
struct App {
    crv::PushConstants pushConstants;
    crv::CameraProperties cameraProperties;
    crv::CloudProperties cloudProperties;
    crv::Parameters parameters;
    crv::ShadowVolumeProperties shadowVolumeProperties;
    float lightDirection[3] = {1, -1, 0};

    std::unique_ptr<crv::Device> device;
    std::unique_ptr<crv::Volume> cloudImage, shadowVolumeImage;
    std::unique_ptr<crv::Image> resultImage;
    std::unique_ptr<crv::RenderTechniqueSV> shadowVolumeTechnique;
    std::unique_ptr<crv::RenderTechniquePT> pathTracingTechnique;
    // --gpus N: this replica's membership of the sample split; its result image then lives in the group's memory
    std::unique_ptr<crv::DeviceGroup> group;
    std::unique_ptr<crv::Stream> stream;

    // main.cpp:223-264
    void UpdateShadowVolume()
    {
        shadowVolumeProperties.SetLightDirection(lightDirection);
        shadowVolumeProperties.SetOrigin(cloudProperties.bounds[0], cloudProperties.bounds[1]);
        crv::DescriptorBufferInfo svInfo{&shadowVolumeProperties, sizeof(shadowVolumeProperties)};
        shadowVolumeTechnique->QueueUpdateShadowVolume(svInfo, 0);
        shadowVolumeTechnique->UpdateDescriptorSets();
        device->WaitIdle();
        shadowVolumeTechnique->RecordDrawCommands(nullptr, 0);
        device->WaitIdle();
        pathTracingTechnique->QueueUpdateShadowVolume(svInfo, 0);
        pathTracingTechnique->UpdateDescriptorSets();
        pushConstants.frameCount = 0;      // headless: zero-cleared image, plain mean
        resultImage->Clear();
    }

    // main.cpp:266-335 (after the grid is in cloudImage)
    void UpdateCloudData()
    {
        crv::DescriptorBufferInfo cloudInfo{&cloudProperties, sizeof(cloudProperties)};
        crv::DescriptorImageInfo cloudImageInfo{cloudImage.get()};
        pathTracingTechnique->QueueUpdateCloudData(cloudInfo, 0);
        pathTracingTechnique->QueueUpdateCloudDataSampler(cloudImageInfo, 0);
        shadowVolumeTechnique->QueueUpdateCloudData(cloudInfo, 0);
        shadowVolumeTechnique->QueueUpdateCloudDataSampler(cloudImageInfo, 0);
        pathTracingTechnique->UpdateDescriptorSets();
        shadowVolumeTechnique->UpdateDescriptorSets();
        UpdateShadowVolume();
    }
};

}  // namespace

int main(int argc, char** argv)
{
    Options o;
    if (!parse(argc, argv, o)) {
        std::fprintf(stderr, "usage: crv_headless [--xyz file | --fbm N] [--width W --height H] [--spp S] [--nsv N] [--g G | --isotropic]\n"
                             "                    [--absorption A] [--density-scaling D] [--light x,y,z] [--fov F] [--per-frame] [--device i] [--pfm out]\n"
                             "                    [--png out.png] [--resume acc.ckpt] [--checkpoint acc.ckpt] [--warmup] [--math strict|driver] [--ppm frames]\n");
        return 2;
    }
    try {
        // one replica of the whole scene per group member (one member = the reference's single-GPU program)
        std::vector<std::unique_ptr<App>> apps;
        std::unique_ptr<crv::Grid3D<float>> grid;
        if (!o.xyz.empty()) {
            grid.reset(crv::Grid3D<float>::Load(o.xyz));
            if (!grid) { std::fprintf(stderr, "crv_headless: cannot open %s\n", o.xyz.c_str()); return 1; }
        }
        for (unsigned m = 0; m < o.gpus; ++m) {
            apps.emplace_back(new App());
            App& app = *apps.back();
            app.device = std::make_unique<crv::Device>(o.same_device ? o.device : o.device + (int)m);
            std::memcpy(app.lightDirection, o.light, sizeof(o.light));
            app.shadowVolumeProperties.voxelAxisCount = o.nsv;
            app.pushConstants.frameCount = 0;
            // SetResolution before SetFOV (main.cpp:383-384)
            app.cameraProperties.SetResolution(o.width, o.height);
            app.cameraProperties.SetFOV(o.fov);
            app.cloudProperties.densityScaling = o.density_scaling;
            if (!o.isotropic) app.parameters.SetPhaseG(o.g);

            // ---- InitializeVulkan (main.cpp:765-910), minus Vulkan ----
            app.shadowVolumeImage = std::make_unique<crv::Volume>(app.device.get(), o.nsv, o.nsv, o.nsv);
            if (o.gpus > 1) {
                app.group = std::make_unique<crv::DeviceGroup>(app.device.get(), m, o.gpus, (uint32_t)o.width, (uint32_t)o.height);
                app.resultImage.reset(new crv::Image(app.group->GetImage(0)));
                app.stream = std::make_unique<crv::Stream>(app.device.get());
            } else {
                app.resultImage = std::make_unique<crv::Image>(app.device.get(), (uint32_t)o.width, (uint32_t)o.height);
            }
            app.shadowVolumeTechnique = std::make_unique<crv::RenderTechniqueSV>(app.device.get(), &app.shadowVolumeProperties, &app.pushConstants);
            app.pathTracingTechnique = std::make_unique<crv::RenderTechniquePT>(app.device.get(), &app.cameraProperties, &app.pushConstants);
            std::vector<crv::FrameImage> svFrames{crv::FrameImage(app.shadowVolumeImage.get())};
            std::vector<crv::FrameImage> ptFrames{crv::FrameImage(app.resultImage.get())};
            app.shadowVolumeTechnique->SetFrameReferences(svFrames);
            app.pathTracingTechnique->SetFrameReferences(ptFrames);
            crv::DescriptorBufferInfo paramInfo{&app.parameters, sizeof(app.parameters)};
            crv::DescriptorBufferInfo cameraInfo{&app.cameraProperties, sizeof(app.cameraProperties)};
            crv::DescriptorImageInfo svImageInfo{app.shadowVolumeImage.get()};
            app.pathTracingTechnique->QueueUpdateParameters(paramInfo, 0);
            app.pathTracingTechnique->QueueUpdateCameraProperties(cameraInfo, 0);
            app.pathTracingTechnique->QueueUpdateShadowVolumeSampler(svImageInfo, 0);
            app.pathTracingTechnique->UpdateDescriptorSets();
            crv::PathTracerExtensions ext;
            ext.absorption = o.absorption;
            ext.math_mode = o.math_driver ? CRV_PT_MATH_DRIVER : CRV_PT_MATH_STRICT;
            app.pathTracingTechnique->SetExtensions(ext);
            app.pathTracingTechnique->SetSkipLevel(o.skip_level);

            // ---- LoadCloudFile (main.cpp:166-190) or the synthetic fBm scene, then SetCloudProperties (main.cpp:144-163) ----
            uint32_t counts[3];
            double voxel[3];
            if (grid) {
                const auto c = grid->GetVoxelCount();
                const auto v = grid->GetVoxelSize();
                for (int i = 0; i < 3; ++i) { counts[i] = c[i]; voxel[i] = v[i]; }
                app.cloudImage = std::make_unique<crv::Volume>(app.device.get(), counts[0], counts[1], counts[2]);
                app.cloudImage->Upload(grid->data());
                app.device->WaitIdle();
            } else {
                for (int i = 0; i < 3; ++i) { counts[i] = o.fbm; voxel[i] = 1.0 / o.fbm; }
                app.cloudImage = std::make_unique<crv::Volume>(app.device.get(), o.fbm, o.fbm, o.fbm);
                app.cloudImage->GenerateFbm(1234);
            }
            app.cloudProperties.SetFromGrid(counts, voxel, app.cloudImage->GetMajorant());
            app.UpdateCloudData();
        }
        if (o.gpus > 1) {        // every member maps everybody else's accumulation image
            std::vector<crv_group_handle> handles;
            for (auto& a : apps) handles.push_back(a->group->Export());
            for (auto& a : apps) a->group->Connect(handles);
        }
        App& app = *apps[0];
        const float sv_ms = app.shadowVolumeTechnique->LastBuildMs();

        // ---- RenderLoop (main.cpp:714-736): srand(0); per frame seed = rand() ----
        // --resume: continue a checkpointed accumulation -- the image, the frameCount the next dispatch pushes, and the
        // rand() sequence from where it stopped (frame f of a run uses element f of srand(0)'s sequence)
        uint32_t first_frame = 0;
        if (!o.resume.empty()) {
            int32_t last_seed = 0;
            app.resultImage->LoadCheckpoint(o.resume.c_str(), &first_frame, &last_seed);
            app.pushConstants.frameCount = first_frame;
        }
        std::vector<int32_t> all_seeds(first_frame + o.spp);
        crv::MsvcRandSequence(0, first_frame + o.spp, all_seeds.data());
        std::vector<int32_t> seeds(all_seeds.begin() + first_frame, all_seeds.end());
        auto render = [&]() {
            if (o.gpus > 1) {
                // frame f -> member f mod N; every member sums its frames, one kernel per member combines and divides.
                // All renders are enqueued before the first combine: a member's combine kernel waits for its peers' on the device.
                // Frames split evenly: every member holds the same number of samples of every pixel, so only the tracked pixels
                // have to cross between the GPUs -- each member's fold stores their sums into every member's inbox and one wait +
                // one local kernel finish the picture (DeviceGroup::PushBegin / PushFinish).  Otherwise whole images are combined.
                const bool even = o.spp % o.gpus == 0;
                for (unsigned m = 0; m < o.gpus; ++m) {
                    App& a = *apps[m];
                    std::vector<int32_t> mine;
                    for (unsigned f = m; f < o.spp; f += o.gpus) mine.push_back(seeds[f]);
                    if (even) {
                        const crv_push_targets targets = a.group->PushBegin();
                        a.pathTracingTechnique->RenderBatchPush(mine.data(), (unsigned)mine.size(), targets, 1.0f / (float)mine.size(), a.stream->Get(), 0);
                    } else {
                        a.resultImage->Clear(a.stream->Get());
                        if (!mine.empty()) a.pathTracingTechnique->RenderBatch(mine.data(), (unsigned)mine.size(), 0, CRV_ACCUM_SUM, a.stream->Get(), 0);
                    }
                }
                for (unsigned m = 0; m < o.gpus; ++m) {
                    App& a = *apps[m];
                    if (even)
                        a.group->PushFinish(1.0f / (float)o.spp, a.pathTracingTechnique->GetPixelLists(), a.stream->Get());
                    else
                        a.group->AllReduceSumScale(1.0f / (float)o.spp, a.stream->Get());
                }
                for (auto& a : apps) a->device->WaitIdle(a->stream->Get());
                for (auto& a : apps) a->group->CheckStatus(a->stream->Get());
                app.pushConstants.frameCount += o.spp;
            } else if (o.per_frame) {
                for (unsigned f = 0; f < o.spp; ++f) {
                    app.pushConstants.seed = seeds[f];
                    app.pathTracingTechnique->RecordDrawCommands(nullptr, 0);
                    app.pushConstants.frameCount++;
                }
            } else {
                app.pathTracingTechnique->RenderBatch(seeds.data(), o.spp, first_frame, CRV_ACCUM_RUNNING_MEAN, nullptr, 0);
                app.pushConstants.frameCount += o.spp;
            }
        };
        if (o.warmup) {
            // --warmup: the same render once untimed (first-use allocations, table builds), then the accumulation is put back
            // where it was -- cleared, or the checkpoint reloaded -- so that the timed run produces the same image
            render();
            for (auto& a : apps) a->device->WaitIdle();
            app.resultImage->Clear();
            app.pushConstants.frameCount = 0;
            if (!o.resume.empty()) {
                int32_t last_seed = 0;
                app.resultImage->LoadCheckpoint(o.resume.c_str(), &first_frame, &last_seed);
                app.pushConstants.frameCount = first_frame;
            }
            for (auto& a : apps) a->pathTracingTechnique->ResetStats();
        }
        for (auto& a : apps) a->device->WaitIdle();
        const auto t0 = std::chrono::steady_clock::now();
        render();
        for (auto& a : apps) a->device->WaitIdle();
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

        std::vector<float> rgba((size_t)o.width * o.height * 4);
        app.resultImage->Download(rgba.data());
        crv_pt_stats st = app.pathTracingTechnique->GetStats();
        // --gpus N: every member must hold the same final picture; the statistics are the members' sums
        bool members_identical = true;
        uint64_t launches = app.device->LaunchCount();
        for (unsigned m = 1; m < o.gpus; ++m) {
            std::vector<float> other(rgba.size());
            apps[m]->resultImage->Download(other.data());
            members_identical = members_identical && std::memcmp(other.data(), rgba.data(), rgba.size() * sizeof(float)) == 0;
            const crv_pt_stats sm = apps[m]->pathTracingTechnique->GetStats();
            st.paths += sm.paths; st.tracking_steps += sm.tracking_steps; st.collisions += sm.collisions; st.tracked_paths += sm.tracked_paths;
            if (!o.same_device) launches += apps[m]->device->LaunchCount();
        }
        if (!o.pfm.empty()) app.resultImage->SavePfm(o.pfm.c_str());                 // linear RGB float
        if (!o.png.empty()) app.resultImage->SavePng(o.png.c_str());                 // what the sRGB swapchain image would show
        if (!o.checkpoint.empty()) app.resultImage->SaveCheckpoint(o.checkpoint.c_str(), app.pushConstants.frameCount, seeds.back());
        // ---- ERenderTechnique::PhotonMapping (main.cpp:204-207, 813): the same scene through RenderTechniquePPM ----
        double ppm_mean = 0.0, ppm_ms = 0.0;
        float ppm_radius = 0.0f;
        if (o.ppm_frames > 0) {
            crv::PhotonMapProperties photonMapProperties;
            photonMapProperties.SetBounds(app.cloudProperties.bounds);                                  // main.cpp:162
            for (int i = 0; i < 3; ++i) photonMapProperties.lightDirection[i] = app.shadowVolumeProperties.lightDirection[i];   // :225
            crv::PushConstants pc;
            pc.frameCount = 0;
            crv::Image ppmImage(app.device.get(), (uint32_t)o.width, (uint32_t)o.height);
            crv::RenderTechniquePPM ppm(app.device.get(), &app.cameraProperties, &photonMapProperties, &pc, 10.0f);
            ppm.AllocateResources();
            std::vector<crv::FrameImage> frames{crv::FrameImage(&ppmImage)};
            ppm.SetFrameReferences(frames);
            crv::DescriptorBufferInfo cloudInfo{&app.cloudProperties, sizeof(app.cloudProperties)};
            crv::DescriptorBufferInfo paramInfo2{&app.parameters, sizeof(app.parameters)};
            crv::DescriptorBufferInfo cameraInfo2{&app.cameraProperties, sizeof(app.cameraProperties)};
            crv::DescriptorBufferInfo svInfo{&app.shadowVolumeProperties, sizeof(app.shadowVolumeProperties)};
            crv::DescriptorImageInfo cloudImageInfo{app.cloudImage.get()};
            crv::DescriptorImageInfo svImageInfo2{app.shadowVolumeImage.get()};
            ppm.QueueUpdateCloudData(cloudInfo, 0);
            ppm.QueueUpdateCloudDataSampler(cloudImageInfo, 0);
            ppm.QueueUpdateParameters(paramInfo2, 0);
            ppm.QueueUpdateCameraProperties(cameraInfo2, 0);
            ppm.QueueUpdateShadowVolume(svInfo, 0);
            ppm.QueueUpdateShadowVolumeSampler(svImageInfo2, 0);
            ppm.UpdateDescriptorSets();
            std::vector<int32_t> ppm_seeds(o.ppm_frames);
            crv::MsvcRandSequence(0, o.ppm_frames, ppm_seeds.data());
            app.device->WaitIdle();
            const auto p0 = std::chrono::steady_clock::now();
            for (unsigned f = 0; f < o.ppm_frames; ++f) {
                pc.seed = ppm_seeds[f];
                ppm.RecordDrawCommands(nullptr, 0);
                pc.frameCount++;
            }
            app.device->WaitIdle();
            ppm_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - p0).count() / o.ppm_frames;
            std::vector<float> px((size_t)o.width * o.height * 4);
            ppmImage.Download(px.data());
            for (float v : px) ppm_mean += v;
            ppm_mean /= (double)px.size();
            ppm_radius = pc.pmRadius;
        }
        std::printf("{\"width\": %d, \"height\": %d, \"spp\": %u, \"mode\": \"%s\", \"gpus\": %u, \"members_identical\": %s, \"render_ms\": %.3f, "
                    "\"mpath_samples_per_s\": %.1f, "
                    "\"sv_build_ms\": %.4f, \"paths\": %llu, \"tracked_paths\": %llu, \"tracking_steps\": %llu, \"collisions\": %llu, \"kernel_launches\": %llu, "
                    "\"image_fnv1a\": \"%016llx\", \"ppm_frames\": %u, \"ppm_frame_ms\": %.3f, \"ppm_image_mean\": %.9g, \"ppm_radius\": %.9g}\n",
                    o.width, o.height, o.spp, o.gpus > 1 ? "split" : o.per_frame ? "per-frame" : "batch", o.gpus, members_identical ? "true" : "false", ms,
                    (double)o.width * o.height * o.spp / ms / 1e3, sv_ms,
                    (unsigned long long)st.paths, (unsigned long long)st.tracked_paths, (unsigned long long)st.tracking_steps, (unsigned long long)st.collisions,
                    (unsigned long long)launches, (unsigned long long)fnv1a(rgba.data(), rgba.size() * sizeof(float)),
                    o.ppm_frames, ppm_ms, ppm_mean, (double)ppm_radius);
        return 0;
    } catch (const crv::Error& e) {
        std::fprintf(stderr, "crv_headless: error %d: %s\n", e.code(), e.what());
        return e.code() == CRV_ERR_NO_DEVICE ? 3 : 1;
    }
}
