// host_runner.cpp — drives the C++ class mirrors (Matcher / CameraGPU / VISystemGPU) on raw arrays written by
// tests/test_gpu_host_classes.py and dumps every public result, so the parity check against the CPU oracle
// lives in pytest next to the other GPU tests.  Usage: host_runner <mode> <dir>
//   mode matcher : d1.bin d2.bin kp1.bin kp2.bin + meta.txt  ->  aux1/aux2/matches/sorted/good dumps
//   mode sequence: frames.bin desc.bin kp.bin rimu.bin tres.bin + meta.txt  ->  poses.bin, ngood.bin, ncand.bin, final.bin
//   mode orb     : frames.bin + meta.txt (frames, w, h, gpu)  ->  per frame orb_kp<i>.bin (x, y, size, angle, response, octave)
//                  and orb_desc<i>.bin, from Camera / CameraGPU::detectAndComputeFeatures with the ORB detector
//   mode nodevice: expects every compute entry to throw vi::DeviceError (run with CUDA_VISIBLE_DEVICES="")
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "vislam/VISystem.hpp"

using namespace cv;
using namespace std;

static map<string, double> read_meta(const string& dir) {
    map<string, double> m;
    ifstream f((dir + "/meta.txt").c_str());
    string k;
    double v;
    while (f >> k >> v) m[k] = v;
    return m;
}

template <typename T>
static vector<T> read_bin(const string& path) {
    ifstream f(path.c_str(), ios::binary | ios::ate);
    if (!f) { cerr << "cannot open " << path << endl; exit(2); }
    const size_t bytes = (size_t)f.tellg();
    f.seekg(0);
    vector<T> v(bytes / sizeof(T));
    f.read(reinterpret_cast<char*>(v.data()), (streamsize)(v.size() * sizeof(T)));
    return v;
}

template <typename T>
static void write_bin(const string& path, const vector<T>& v) {
    ofstream f(path.c_str(), ios::binary);
    f.write(reinterpret_cast<const char*>(v.data()), (streamsize)(v.size() * sizeof(T)));
}

static void dump_dmatch(const string& path, const vector<DMatch>& m) {
    vector<float> o;
    for (size_t i = 0; i < m.size(); i++) {
        o.push_back((float)m[i].queryIdx); o.push_back((float)m[i].trainIdx);
        o.push_back((float)m[i].imgIdx); o.push_back(m[i].distance);
    }
    write_bin(path, o);
}

static void dump_knn(const string& path, const vector<vector<DMatch> >& m) {
    vector<float> o;   // per row: size, then (queryIdx, trainIdx, imgIdx, distance) x 2 (zero padded)
    for (size_t i = 0; i < m.size(); i++) {
        o.push_back((float)m[i].size());
        for (size_t k = 0; k < 2; k++) {
            if (k < m[i].size()) {
                o.push_back((float)m[i][k].queryIdx); o.push_back((float)m[i][k].trainIdx);
                o.push_back((float)m[i][k].imgIdx); o.push_back(m[i][k].distance);
            } else {
                for (int z = 0; z < 4; z++) o.push_back(0.f);
            }
        }
    }
    write_bin(path, o);
}

static vector<KeyPoint> to_keypoints(const float* xy, int n) {
    vector<KeyPoint> k((size_t)n);
    for (int i = 0; i < n; i++) { k[i].pt.x = xy[2 * i]; k[i].pt.y = xy[2 * i + 1]; }
    return k;
}

static int run_matcher(const string& dir) {
    map<string, double> meta = read_meta(dir);
    const int n1 = (int)meta["n1"], n2 = (int)meta["n2"], dim = (int)meta["dim"], norm = (int)meta["norm"];
    const int w = (int)meta["w"], h = (int)meta["h"], n_cells = (int)meta["n_cells"], sym_mode = (int)meta["sym_mode"];
    vector<float> kp1 = read_bin<float>(dir + "/kp1.bin"), kp2 = read_bin<float>(dir + "/kp2.bin");
    Mat d1, d2;
    vector<uint8_t> b1, b2;
    vector<float> f1, f2;
    if (norm == 1) {
        b1 = read_bin<uint8_t>(dir + "/d1.bin"); b2 = read_bin<uint8_t>(dir + "/d2.bin");
        d1 = Mat(n1, dim, CV_8U, b1.data()); d2 = Mat(n2, dim, CV_8U, b2.data());
    } else {
        f1 = read_bin<float>(dir + "/d1.bin"); f2 = read_bin<float>(dir + "/d2.bin");
        d1 = Mat(n1, dim, CV_32F, f1.data()); d2 = Mat(n2, dim, CV_32F, f2.data());
    }
    MatcherGPU m(norm == 1 ? USE_BRUTE_FORCE_GPU_HAMMING : USE_BRUTE_FORCE_GPU);
    m.sym_mode = sym_mode;
    m.setImageDimensions(w, h);
    m.clear();
    m.setKeypoints(to_keypoints(kp1.data(), n1), to_keypoints(kp2.data(), n2));
    m.setDescriptors(d1, d2);
    m.computeGPUMatches();
    dump_knn(dir + "/aux1_raw.bin", m.aux_matches1);
    dump_knn(dir + "/aux2_raw.bin", m.aux_matches2);
    m.computeBestMatches(n_cells);
    dump_knn(dir + "/aux1_filtered.bin", m.aux_matches1);
    dump_knn(dir + "/aux2_filtered.bin", m.aux_matches2);
    dump_dmatch(dir + "/matches.bin", m.matches);
    dump_dmatch(dir + "/sorted.bin", m.sortedMatches);
    dump_dmatch(dir + "/good.bin", m.goodMatches);
    vector<KeyPoint> g1, g2;
    m.getGoodMatches(g1, g2);
    vector<float> gk;
    for (size_t i = 0; i < g1.size(); i++) { gk.push_back(g1[i].pt.x); gk.push_back(g1[i].pt.y); gk.push_back(g2[i].pt.x); gk.push_back(g2[i].pt.y); }
    write_bin(dir + "/good_kp.bin", gk);
    vector<float> counts;
    counts.push_back((float)m.nSymMatches); counts.push_back((float)m.nBestMatches);
    write_bin(dir + "/counts.bin", counts);
    cout << "matcher ok: sym " << m.nSymMatches << " good " << m.nBestMatches << " launches " << vi::Device::get().launches() << endl;
    return 0;
}

static int run_sequence(const string& dir) {
    map<string, double> meta = read_meta(dir);
    const int T = (int)meta["frames"], N = (int)meta["n_feat"], w = (int)meta["w"], h = (int)meta["h"];
    const int n_cells = (int)meta["n_cells"], mirror = (int)meta["mirror_host"], grad_images = (int)meta["grad_images"];
    const bool orb = meta.count("orb") && meta["orb"] != 0;      // features from the device ORB detector instead of files
    vector<uint8_t> frames = read_bin<uint8_t>(dir + "/frames.bin"), desc;
    vector<float> kp, rimu = read_bin<float>(dir + "/rimu.bin"), tres = read_bin<float>(dir + "/tres.bin");
    if (!orb) { desc = read_bin<uint8_t>(dir + "/desc.bin"); kp = read_bin<float>(dir + "/kp.bin"); }
    vi::VISystemGPU sys;
    Mat Kmat = Mat::zeros(3, 3, CV_32F);
    Kmat.at<float>(0, 0) = (float)meta["fx"]; Kmat.at<float>(1, 1) = (float)meta["fy"];
    Kmat.at<float>(0, 2) = (float)meta["cx"]; Kmat.at<float>(1, 2) = (float)meta["cy"]; Kmat.at<float>(2, 2) = 1.f;
    sys.InitializePyramid(w, h, Kmat);
    sys.InitializeCameraGPU(USE_ORB, USE_BRUTE_FORCE_GPU_HAMMING, w, h, n_cells, 7);
    sys.cameraGPU.mirror_host = mirror != 0;
    sys.track_from_estimate = true;
    sys.keep_trace = true;
    vector<float> poses, finals, ngood, ncand, niter;
    for (int t = 0; t < T; t++) {
        Mat img(h, w, CV_8U, frames.data() + (size_t)t * w * h);
        const float* kxy = orb ? nullptr : kp.data() + (size_t)t * N * 2;
        uint8_t* dd = orb ? nullptr : desc.data() + (size_t)t * N * 32;
        if (!orb)
            sys.cameraGPU.featureProvider = [&](const Mat&, vector<KeyPoint>& k, Mat& d) {
                k = to_keypoints(kxy, N);
                d = Mat(N, 32, CV_8U, dd);
            };
        if (t > 0) {
            const float* r = rimu.data() + (size_t)(t - 1) * 9;
            sys.RotationResidualImu = Matx33f(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]);
            Mat tr(3, 1, CV_32F);
            for (int i = 0; i < 3; i++) tr.at<float>(i, 0) = tres[(size_t)(t - 1) * 3 + i];
            sys.setGtRes(tr, Mat());
        }
        if (!grad_images && t > 0) {
            // exercise the on-the-fly gradient path: drop the previous frame's gradient images
            Frame* p = sys.cameraGPU.frameList.back();
            p->grad_on_device = false;
        }
        sys.AddFrameGPU(img, vector<Point3d>(), vector<Point3d>());
        if (t > 0) {
            Frame* prev = sys.cameraGPU.frameList[sys.cameraGPU.frameList.size() - 2];
            for (int i = 0; i < 7; i++) poses.push_back(prev->rigid_transformation_.data()[i]);
            for (int i = 0; i < 7; i++) finals.push_back(sys.final_poseCam.data()[i]);
            ngood.push_back((float)prev->nextGoodMatches.size());
            for (int l = 0; l < 5; l++) ncand.push_back((float)prev->n_cand[l]);
            niter.push_back((float)sys.last_trace.size());
            if (mirror) {
                for (int l = 0; l < 5; l++)
                    if (prev->candidatePoints[l].rows != prev->n_cand[l]) { cerr << "candidate mirror mismatch" << endl; return 3; }
                if (prev->gradientX[0].empty() || prev->grayImage[4].empty()) { cerr << "host mirrors missing" << endl; return 3; }
            }
        }
    }
    write_bin(dir + "/poses.bin", poses);
    write_bin(dir + "/final.bin", finals);
    write_bin(dir + "/ngood.bin", ngood);
    write_bin(dir + "/ncand.bin", ncand);
    write_bin(dir + "/niter.bin", niter);
    // the last frame's host mirrors, for a spot check of Camera::Update / computeGradient against the oracle
    Frame* last = sys.cameraGPU.frameList.back();
    if (mirror) {
        vector<uint8_t> g4(last->grayImage[4].data, last->grayImage[4].data + last->grayImage[4].total());
        write_bin(dir + "/last_gray4.bin", g4);
        vector<int16_t> gx3(last->gradientX[3].ptr<int16_t>(), last->gradientX[3].ptr<int16_t>() + last->gradientX[3].total());
        write_bin(dir + "/last_gx3.bin", gx3);
    }
    // WarpFunctionSE3 on its own
    Frame* prev = sys.cameraGPU.frameList[sys.cameraGPU.frameList.size() - 2];
    if (mirror && prev->candidatePoints[2].rows > 0) {
        Mat wp = sys.WarpFunctionSE3(prev->candidatePoints[2], prev->rigid_transformation_, 2);
        vector<float> o(wp.ptr<float>(), wp.ptr<float>() + wp.total());
        write_bin(dir + "/warp2.bin", o);
        vector<float> in(prev->candidatePoints[2].ptr<float>(), prev->candidatePoints[2].ptr<float>() + prev->candidatePoints[2].total());
        write_bin(dir + "/warp2_in.bin", in);
    }
    cout << "sequence ok: " << T - 1 << " pairs, launches " << vi::Device::get().launches() << endl;
    return 0;
}

static int run_nodevice() {
    int thrown = 0;
    try {
        Matcher m(USE_BRUTE_FORCE_HAMMING);
        vector<uint8_t> b(64, 1);
        m.setDescriptors(Mat(2, 32, CV_8U, b.data()), Mat(2, 32, CV_8U, b.data()));
        m.computeMatches();
    } catch (const vi::DeviceError& e) {
        thrown++;
        cout << "computeMatches: " << e.what() << endl;
    }
    try {
        Camera c(USE_ORB, USE_BRUTE_FORCE_HAMMING, 64, 48, 49, 7);
        vector<uint8_t> img(64 * 48, 7);
        c.Update(Mat(48, 64, CV_8U, img.data()));
    } catch (const vi::DeviceError& e) {
        thrown++;
        cout << "Update: " << e.what() << endl;
    }
    cout << "nodevice: " << thrown << " of 2 entries failed loudly" << endl;
    return thrown == 2 ? 0 : 4;
}

static int run_orb(const string& dir) {
    map<string, double> meta = read_meta(dir);
    const int T = (int)meta["frames"], w = (int)meta["w"], h = (int)meta["h"];
    const bool gpu = meta["gpu"] != 0;
    vector<uint8_t> frames = read_bin<uint8_t>(dir + "/frames.bin");
    Camera cam;
    CameraGPU camg;
    if (gpu) camg.initializateCameraGPU(USE_ORB, USE_BRUTE_FORCE_HAMMING, w, h, 49, 5);
    else cam.initializate(USE_ORB, USE_BRUTE_FORCE_HAMMING, w, h, 49, 5);
    Camera& c = gpu ? static_cast<Camera&>(camg) : cam;
    for (int i = 0; i < T; i++) {
        Mat img(h, w, CV_8U);
        memcpy(img.data, frames.data() + (size_t)i * w * h, (size_t)w * h);
        c.Update(img);
        const int n = gpu ? camg.detectAndComputeGPUFeatures() : c.detectAndComputeFeatures();
        vector<float> kp;
        for (int k = 0; k < n; k++) {
            const KeyPoint& q = c.currentFrame->keypoints[k];
            kp.push_back(q.pt.x); kp.push_back(q.pt.y); kp.push_back(q.size); kp.push_back(q.angle); kp.push_back(q.response);
            kp.push_back((float)q.octave);
        }
        ostringstream a, b;
        a << dir << "/orb_kp" << i << ".bin";
        b << dir << "/orb_desc" << i << ".bin";
        write_bin(a.str(), kp);
        vector<uint8_t> d((size_t)n * 32);
        if (n) memcpy(d.data(), c.currentFrame->descriptors.data, d.size());
        write_bin(b.str(), d);
        if (c.currentFrame->descriptors.rows != n || (n && c.currentFrame->descriptors.cols != 32)) return 3;
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { cerr << "usage: host_runner <matcher|sequence|orb|nodevice> [dir]" << endl; return 1; }
    const string mode = argv[1];
    try {
        if (mode == "nodevice") return run_nodevice();
        if (argc < 3) return 1;
        if (mode == "matcher") return run_matcher(argv[2]);
        if (mode == "sequence") return run_sequence(argv[2]);
        if (mode == "orb") return run_orb(argv[2]);
    } catch (const std::exception& e) {
        cerr << "host_runner: exception: " << e.what() << endl;
        return 5;
    }
    return 1;
}
