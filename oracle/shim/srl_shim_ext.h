// TEST INFRASTRUCTURE (oracle/): declaration-only stand-ins for ROS, PCL, OpenCV, tf, Ceres/glog and the message packages,
// just enough for the reference's HEADERS (include/*.h) to parse when src/{optimize,utility,eskfEstimator,cloudMap,state}.cpp
// are compiled where they lie (oracle/Makefile, target _ref/libsrl_reference.so).  Nothing here computes anything the
// scan-matching path uses; the reference's ROS node, sensor decoding and vision module are neither compiled nor linked.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <stdexcept>
#include <cmath>
#include <sstream>
#include <string>
#include <vector>
#include <Eigen/Core>

typedef unsigned char uchar;

// ---- glog / rosconsole
struct SrlNullStream { template <class T> SrlNullStream& operator<<(const T&) { return *this; } SrlNullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } };
#define LOG(severity) SrlNullStream()
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)

// ---- ROS
namespace ros {
struct Time { double t = 0; Time() {} double toSec() const { return t; } Time& fromSec(double s) { t = s; return *this; } static Time now() { return Time(); } static void init() {} };
struct Duration { Duration(double = 0) {} void sleep() {} };
struct Rate { Rate(double) {} void sleep() {} };
struct Publisher { template <class M> void publish(const M&) const {} int getNumSubscribers() const { return 0; } };
struct Subscriber {};
struct NodeHandle {
    NodeHandle() {}
    NodeHandle(const std::string&) {}
    // no parameter server: every parameter takes the default the reference passes; the array parameters whose reference
    // default is empty (and which the constructor then indexes) get neutral values of the right length
    template <class T, class U> bool param(const std::string&, T& v, const U& d) { v = d; return false; }
    template <class T> bool param(const std::string& name, std::vector<double>& v, const std::vector<double>&) {
        auto has = [&](const char* k) { return name.find(k) != std::string::npos; };
        if (has("_R_") || has("camera_intrinsic")) v = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        else if (has("dist_coeffs")) v = {0, 0, 0, 0, 0};
        else if (has("gravity")) v = {0, 0, 9.81};
        else v = {0, 0, 0};
        return false;
    }
    template <class M, class... A> Publisher advertise(A&&...) { return Publisher(); }
    template <class M, class... A> Subscriber subscribe(const std::string&, int, A&&...) { return Subscriber(); }
    template <class C, class Arg> Subscriber subscribe(const std::string&, int, void (C::*)(Arg), C*) { return Subscriber(); }
};
inline bool ok() { return false; }
inline void spinOnce() {}
inline void spin() {}
template <class... A> inline void init(A&&...) {}
}  // namespace ros
namespace std_msgs { struct Header { ros::Time stamp; std::string frame_id; unsigned seq = 0; }; }
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct PoseWithCovariance { Pose pose; double covariance[36]; };
struct Twist { Vector3 linear, angular; };
struct TwistWithCovariance { Twist twist; double covariance[36]; };
}  // namespace geometry_msgs
namespace nav_msgs {
struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; };
struct Odometry { std_msgs::Header header; std::string child_frame_id; geometry_msgs::PoseWithCovariance pose; geometry_msgs::TwistWithCovariance twist; };
}  // namespace nav_msgs
namespace sensor_msgs {
struct Imu { std_msgs::Header header; geometry_msgs::Quaternion orientation; geometry_msgs::Vector3 angular_velocity, linear_acceleration;
             typedef std::shared_ptr<const Imu> ConstPtr; typedef std::shared_ptr<Imu> Ptr; };
typedef std::shared_ptr<const Imu> ImuConstPtr;
struct PointField { std::string name; unsigned offset = 0; unsigned char datatype = 0; unsigned count = 0; };
struct PointCloud2 { std_msgs::Header header; std::vector<PointField> fields; std::vector<unsigned char> data; unsigned width = 0, height = 0;
                     typedef std::shared_ptr<const PointCloud2> ConstPtr; typedef std::shared_ptr<PointCloud2> Ptr; };
struct Image { std_msgs::Header header; typedef std::shared_ptr<const Image> ConstPtr; };
namespace image_encodings { static const char* const BGR8 = "bgr8"; static const char* const RGB8 = "rgb8"; static const char* const MONO8 = "mono8"; }
typedef std::shared_ptr<const Image> ImageConstPtr;
struct CompressedImage { std_msgs::Header header; std::string format; typedef std::shared_ptr<const CompressedImage> ConstPtr; };
typedef std::shared_ptr<const CompressedImage> CompressedImageConstPtr;
}  // namespace sensor_msgs
namespace livox_ros_driver {
struct CustomPoint { unsigned offset_time = 0; float x = 0, y = 0, z = 0; unsigned char reflectivity = 0, tag = 0, line = 0; };
struct CustomMsg { std_msgs::Header header; uint64_t timebase = 0; unsigned point_num = 0; unsigned char lidar_id = 0; std::vector<CustomPoint> points;
                   typedef std::shared_ptr<const CustomMsg> ConstPtr; };
}  // namespace livox_ros_driver
namespace tf {
struct Quaternion { double x_ = 0, y_ = 0, z_ = 0, w_ = 1; Quaternion() {} Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {} };
struct Vector3 { double v[3]; Vector3() {} Vector3(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; } };
struct Transform { void setRotation(const Quaternion&) {} void setOrigin(const Vector3&) {} };
struct StampedTransform : Transform { ros::Time stamp_; std::string frame_id_, child_frame_id_; };
inline geometry_msgs::Quaternion createQuaternionMsgFromRollPitchYaw(double, double, double) { return geometry_msgs::Quaternion(); }
struct TransformBroadcaster { template <class T> void sendTransform(const T&) {} };
}  // namespace tf

// ---- PCL
#define PCL_ADD_POINT4D union { float data[4]; struct { float x; float y; float z; }; }
#define POINT_CLOUD_REGISTER_POINT_STRUCT(...)
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
struct PointXYZRGB { float x = 0, y = 0, z = 0; unsigned char r = 0, g = 0, b = 0, a = 0; };
struct PointXYZINormal { float x = 0, y = 0, z = 0, intensity = 0, normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
template <class P> struct PointCloud {
    typedef std::shared_ptr<PointCloud<P>> Ptr;
    typedef std::shared_ptr<const PointCloud<P>> ConstPtr;
    std::vector<P> points; unsigned width = 0, height = 1; bool is_dense = true;
    size_t size() const { return points.size(); } void clear() { points.clear(); } void push_back(const P& p) { points.push_back(p); }
    void resize(size_t n) { points.resize(n); } void reserve(size_t n) { points.reserve(n); } P& operator[](size_t i) { return points[i]; }
};
template <class C> inline void toROSMsg(const C&, sensor_msgs::PointCloud2&) {}
namespace io { template <class C> inline int savePCDFileBinary(const std::string&, const C&) { return 0; } }
struct PCDWriter { template <class C> int writeBinary(const std::string&, const C&) { return 0; } };
template <class P> struct VoxelGrid {};
}  // namespace pcl

// ---- OpenCV (types that appear in class declarations and inline getters only)
namespace cv {
template <class T> struct Point_ { T x = 0, y = 0; Point_() {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f; typedef Point_<double> Point2d; typedef Point_<int> Point;
template <class T> struct Size_ { T width = 0, height = 0; Size_() {} Size_(T w, T h) : width(w), height(h) {} };
typedef Size_<int> Size;
struct Range { int start = 0, end = 0; Range() {} Range(int s, int e) : start(s), end(e) {} };
// cv::saturate_cast (core/saturate.hpp): double -> uchar goes through cvRound (lrint: round half to even), then clamps
template <class T> inline T saturate_cast(double v) { return static_cast<T>(v); }
template <class T> inline T saturate_cast(int v) { return static_cast<T>(v); }
template <class T> inline T saturate_cast(float v) { return static_cast<T>(v); }
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255u ? v : v > 0 ? 255 : 0); }
template <> inline uchar saturate_cast<uchar>(double v) { return saturate_cast<uchar>((int)lrint(v)); }
template <> inline uchar saturate_cast<uchar>(float v) { return saturate_cast<uchar>((int)lrintf(v)); }
// cv::Vec with Matx's element-wise saturating arithmetic (core/matx.hpp: Matx_AddOp, Matx_SubOp, Matx_ScaleOp)
template <class T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N == 3, "3 channels"); val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; } const T& operator[](int i) const { return val[i]; }
    T& operator()(int i) { return val[i]; } const T& operator()(int i) const { return val[i]; }
};
template <class T, int N> inline Vec<T, N> operator+(const Vec<T, N>& a, const Vec<T, N>& b) { Vec<T, N> r; for (int i = 0; i < N; ++i) r.val[i] = saturate_cast<T>(a.val[i] + b.val[i]); return r; }
template <class T, int N> inline Vec<T, N> operator-(const Vec<T, N>& a, const Vec<T, N>& b) { Vec<T, N> r; for (int i = 0; i < N; ++i) r.val[i] = saturate_cast<T>(a.val[i] - b.val[i]); return r; }
template <class T, int N> inline Vec<T, N> operator*(double alpha, const Vec<T, N>& a) { Vec<T, N> r; for (int i = 0; i < N; ++i) r.val[i] = saturate_cast<T>(a.val[i] * alpha); return r; }
template <class T, int N> inline Vec<T, N> operator*(const Vec<T, N>& a, double alpha) { return alpha * a; }
template <class T1, class T2, int N> inline Vec<T1, N>& operator+=(Vec<T1, N>& a, const Vec<T2, N>& b) { for (int i = 0; i < N; ++i) a.val[i] = saturate_cast<T1>(a.val[i] + b.val[i]); return a; }
typedef Vec<uchar, 3> Vec3b; typedef Vec<float, 3> Vec3f; typedef Vec<double, 3> Vec3d;
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };
// cv::Mat: a shared, row-major byte image (what cloudFrame::getRgb / getSubPixel read through ptr<T>(row) and at<T>(row, col))
struct Mat {
    int rows = 0, cols = 0; size_t step = 0; std::shared_ptr<std::vector<uchar>> buf; uchar* data = nullptr;
    Mat() {}
    Mat(int r, int c, size_t elem) { create(r, c, elem); }
    void create(int r, int c, size_t elem) { rows = r; cols = c; step = (size_t)c * elem; buf = std::make_shared<std::vector<uchar>>((size_t)r * step); data = buf->data(); }
    Mat clone() const { Mat m(*this); if (buf) { m.buf = std::make_shared<std::vector<uchar>>(*buf); m.data = m.buf->data(); } return m; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
};
struct TermCriteria { enum { COUNT = 1, MAX_ITER = 1, EPS = 2 }; int type = 0, maxCount = 0; double epsilon = 0;
                      TermCriteria() {} TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {} };
struct _InputArray { _InputArray() {} template <class T> _InputArray(const T&) {} };
typedef const _InputArray& InputArray; typedef const _InputArray& OutputArray; typedef const _InputArray& InputOutputArray;
typedef const _InputArray& OutputArrayOfArrays; typedef const _InputArray& InputArrayOfArrays;
inline const _InputArray& noArray() { static _InputArray a; return a; }
struct ParallelLoopBody { virtual ~ParallelLoopBody() {} virtual void operator()(const Range&) const = 0; };
template <class F> inline void parallel_for_(const Range& r, const F& f, double = -1.) { f(r); }   // sequential: deterministic
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4 };
struct RNG { uint64_t state; RNG(uint64_t s = 0xffffffff) : state(s) {} int uniform(int a, int) { return a; } double uniform(double a, double) { return a; } };
template <class T> struct Ptr : std::shared_ptr<T> { using std::shared_ptr<T>::shared_ptr; };
}  // namespace cv
namespace cv_bridge {
struct CvImage { cv::Mat image; }; typedef std::shared_ptr<CvImage> CvImagePtr;
struct Exception : std::runtime_error { Exception() : std::runtime_error("cv_bridge") {} };
template <class M> inline CvImagePtr toCvCopy(const M&, const char* = "") { return std::make_shared<CvImage>(); }
}
namespace boost { typedef std::mt19937_64 mt19937_64; }
