// kiss_icp_pybind for the B200 backend: the module the reference's Python package imports as
// `kiss_icp.pybind.kiss_icp_pybind` (python/kiss_icp/pybind/kiss_icp_pybind.cpp:44-144), re-bound over the C-ABI of
// include/kiss_icp_b200.h. Same names, same keyword arguments, same exceptions; needs neither Eigen nor Sophus nor TBB.
// With this module on its path, python/kiss_icp/*.py runs unchanged on the GPU library.
//
//   g++ -O2 -std=c++17 -shared -fPIC $(python -m pybind11 --includes) bindings/kiss_icp_pybind.cpp \
//       -Iinclude -Lkiss-icp_b200 -lkiss_icp_b200 -Wl,-rpath,'$ORIGIN/../kiss-icp_b200' \
//       -o bindings/kiss_icp_pybind$(python3-config --extension-suffix)
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <stdexcept>

#include "kiss_icp_b200.h"

namespace py = pybind11;
using namespace py::literals;
using arr = py::array_t<double, py::array::c_style | py::array::forcecast>;

namespace {
// kb_status -> the exceptions the reference raised through pybind11
void check(int st) {
    if (st == KB_OK) return;
    if (st == KB_ERR_OUT_OF_RANGE) throw py::index_error(kb_last_error());  // std::out_of_range (Preprocessing.cpp:76-77)
    if (st == KB_ERR_NOT_SE3) throw py::value_error(kb_last_error());       // Sophus ENSURE (the reference aborts)
    throw std::runtime_error(kb_last_error());
}
// std::vector<Eigen::Vector3d> arguments: (N,3) float64, like py_array_to_vectors_double (stl_vector_eigen.h:67-80)
const arr &points_arg(const arr &a) {
    if (a.ndim() != 2 || a.shape(1) != 3) throw py::cast_error("expected an (N, 3) array");
    return a;
}
const double *mat4_arg(const arr &a) {  // Eigen::Matrix4d arguments: NumPy is row-major, so is the C-ABI
    if (a.ndim() != 2 || a.shape(0) != 4 || a.shape(1) != 4) throw py::cast_error("expected a (4, 4) array");
    return a.data();
}
// blocking GPU work runs without the GIL (other Python threads — a ROS executor, a visualizer — keep running; the
// reference holds it, SURVEY.md 8b "Threading")
template <class F>
void nogil(F &&call) {
    int st;
    {
        py::gil_scoped_release release;
        st = call();
    }
    check(st);
}
arr points_out(size_t n) { return arr({static_cast<py::ssize_t>(n), static_cast<py::ssize_t>(3)}); }
py::object first_rows(const arr &a, size_t n) {  // the first n points of an output buffer sized for the worst case
    return a[py::slice(0, static_cast<py::ssize_t>(n), 1)].attr("copy")();
}

struct Map {  // _VoxelHashMap
    kb_map *h = nullptr;
    Map(double voxel_size, double max_distance, int max_points_per_voxel) {
        check(kb_map_create(voxel_size, max_distance, static_cast<unsigned>(max_points_per_voxel), &h));
    }
    ~Map() { kb_map_destroy(h); }
    Map(const Map &) = delete;
    Map &operator=(const Map &) = delete;
};
struct Pre {  // _Preprocessor
    kb_preprocessor *h = nullptr;
    Pre(double max_range, double min_range, bool deskew, int max_num_threads) {
        check(kb_preprocessor_create(max_range, min_range, deskew ? 1 : 0, max_num_threads, &h));
    }
    ~Pre() { kb_preprocessor_destroy(h); }
    Pre(const Pre &) = delete;
    Pre &operator=(const Pre &) = delete;
};
struct Reg {  // _Registration
    kb_registration *h = nullptr;
    Reg(int max_num_iterations, double convergence_criterion, int max_num_threads) {
        check(kb_registration_create(max_num_iterations, convergence_criterion, max_num_threads, &h));
    }
    ~Reg() { kb_registration_destroy(h); }
    Reg(const Reg &) = delete;
    Reg &operator=(const Reg &) = delete;
};
struct Thr {  // _AdaptiveThreshold
    kb_threshold *h = nullptr;
    Thr(double initial_threshold, double min_motion_th, double max_range) {
        check(kb_threshold_create(initial_threshold, min_motion_th, max_range, &h));
    }
    ~Thr() { kb_threshold_destroy(h); }
    Thr(const Thr &) = delete;
    Thr &operator=(const Thr &) = delete;
};
}  // namespace

PYBIND11_MODULE(kiss_icp_pybind, m) {
    m.doc() = "kiss_icp_pybind over libkiss_icp_b200 (B200 backend)";
    // `_Vector3dVector(points)`: the reference wraps std::vector<Eigen::Vector3d> (opaque, buffer protocol). A C-contiguous
    // (N,3) float64 ndarray has the same memory layout and already offers len(), bool-ness via len, copy and np.asarray,
    // so here the "vector" IS that array.
    m.def("_Vector3dVector", [](const arr &a) { return arr(points_arg(a)); });

    py::class_<Map>(m, "_VoxelHashMap", "Don't use this")
        .def(py::init<double, double, int>(), "voxel_size"_a, "max_distance"_a, "max_points_per_voxel"_a)
        .def("_clear", [](Map &s) { check(kb_map_clear(s.h)); })
        .def("_empty",
             [](Map &s) {
                 int e = 1;
                 check(kb_map_empty(s.h, &e));
                 return e != 0;
             })
        // the two _update overloads of the reference (kiss_icp_pybind.cpp:59-70): (points, origin[3]) and (points, pose[4,4]),
        // told apart like pybind11/eigen.h tells Vector3d from Matrix4d: by shape, next overload on a mismatch
        .def(
            "_update",
            [](Map &s, const arr &pts, const arr &origin) {
                if (origin.ndim() != 1 || origin.shape(0) != 3) throw py::reference_cast_error();  // -> try (points, pose)
                points_arg(pts);
                nogil([&] { return kb_map_update_origin(s.h, pts.data(), pts.shape(0), origin.data()); });
            },
            "points"_a, "origin"_a)
        .def(
            "_update",
            [](Map &s, const arr &pts, const arr &pose) {
                if (pose.ndim() != 2 || pose.shape(0) != 4 || pose.shape(1) != 4) throw py::reference_cast_error();  // -> TypeError
                points_arg(pts);
                nogil([&] { return kb_map_update_pose(s.h, pts.data(), pts.shape(0), pose.data()); });
            },
            "points"_a, "pose"_a)
        .def(
            "_add_points", [](Map &s, const arr &pts) { points_arg(pts); nogil([&] { return kb_map_add_points(s.h, pts.data(), pts.shape(0)); }); },
            "points"_a)
        .def(
            "_remove_far_away_points",
            [](Map &s, const arr &origin) {
                if (origin.ndim() != 1 || origin.shape(0) != 3) throw py::cast_error("expected an origin (3,)");
                nogil([&] { return kb_map_remove_far(s.h, origin.data()); });
            },
            "origin"_a)
        .def("_point_cloud", [](Map &s) {
            size_t n = 0;
            check(kb_map_pointcloud(s.h, nullptr, 0, &n));
            arr out = points_out(n);
            if (n) check(kb_map_pointcloud(s.h, out.mutable_data(), n, &n));
            return out;
        });

    py::class_<Pre>(m, "_Preprocessor", "Don't use this")
        .def(py::init<double, double, bool, int>(), "max_range"_a, "min_range"_a, "deskew"_a, "max_num_threads"_a)
        .def(
            "_preprocess",
            [](Pre &s, const arr &pts, const arr &timestamps, const arr &relative_motion) {
                points_arg(pts);
                const size_t n = static_cast<size_t>(pts.shape(0));
                arr out = points_out(n);
                size_t kept = 0;
                const double *M = mat4_arg(relative_motion);
                double *o = out.mutable_data();
                nogil([&] { return kb_preprocessor_preprocess(s.h, pts.data(), n, timestamps.data(), static_cast<size_t>(timestamps.size()), M, o, n, &kept); });
                return first_rows(out, kept);
            },
            "points"_a, "timestamps"_a, "relative_motion"_a);

    py::class_<Reg>(m, "_Registration", "Don't use this")
        .def(py::init<int, double, int>(), "max_num_iterations"_a, "convergence_criterion"_a, "max_num_threads"_a)
        .def(
            "_align_points_to_map",
            [](Reg &s, const arr &pts, const Map &voxel_map, const arr &initial_guess, double max_correspondance_distance,
               double kernel) {
                points_arg(pts);
                arr out({static_cast<py::ssize_t>(4), static_cast<py::ssize_t>(4)});
                const double *G = mat4_arg(initial_guess);
                double *o = out.mutable_data();
                nogil([&] { return kb_registration_align_points_to_map(s.h, pts.data(), pts.shape(0), voxel_map.h, G, max_correspondance_distance, kernel, o); });
                return out;
            },
            "points"_a, "voxel_map"_a, "initial_guess"_a, "max_correspondance_distance"_a, "kernel"_a);

    py::class_<Thr>(m, "_AdaptiveThreshold", "Don't use this")
        .def(py::init<double, double, double>(), "initial_threshold"_a, "min_motion_th"_a, "max_range"_a)
        .def("_compute_threshold",
             [](Thr &s) {
                 double sigma = 0.0;
                 check(kb_threshold_compute(s.h, &sigma));
                 return sigma;
             })
        .def(
            "_update_model_deviation",
            [](Thr &s, const arr &model_deviation) { check(kb_threshold_update_model_deviation(s.h, mat4_arg(model_deviation))); },
            "model_deviation"_a);

    m.def(
        "_voxel_down_sample",
        [](const arr &frame, double voxel_size) {
            points_arg(frame);
            const size_t n = static_cast<size_t>(frame.shape(0));
            arr out = points_out(n);
            size_t kept = 0;
            double *o = out.mutable_data();
            nogil([&] { return kb_voxel_down_sample(frame.data(), n, voxel_size, o, n, &kept); });
            return first_rows(out, kept);
        },
        "frame"_a, "voxel_size"_a);
    m.def(
        "_correct_kitti_scan",
        [](const arr &frame) {
            points_arg(frame);
            arr out = points_out(static_cast<size_t>(frame.shape(0)));
            check(kb_correct_kitti_scan(frame.data(), static_cast<size_t>(frame.shape(0)), out.mutable_data()));
            return out;
        },
        "frame"_a);
    // Metrics stay host code (cpp/kiss_icp/metrics/Metrics.cpp); this backend keeps them in kiss_icp_b200/metrics.py
    m.def(
        "_kitti_seq_error",
        [](const py::object &gt_poses, const py::object &results_poses) {
            return py::module_::import("kiss_icp_b200.metrics").attr("sequence_error")(gt_poses, results_poses);
        },
        "gt_poses"_a, "results_poses"_a);
    m.def(
        "_absolute_trajectory_error",
        [](const py::object &gt_poses, const py::object &results_poses) {
            return py::module_::import("kiss_icp_b200.metrics").attr("absolute_trajectory_error")(gt_poses, results_poses);
        },
        "gt_poses"_a, "results_poses"_a);
}
