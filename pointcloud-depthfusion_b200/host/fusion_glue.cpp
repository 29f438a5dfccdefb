/*
 * fusion_glue.cpp — the host-side matrix / intrinsics glue FusionNode does around the hot path,
 * without Eigen (absent here; the reference uses Eigen 3.4.0, README.md:50). Pure host double math,
 * callable without a GPU. The Eigen overloads in include/pointcloud_processing/pointcloud.h forward
 * here when Eigen is available to the consumer.
 *
 *   dfb_interpolate_transform      FusionNode::interpolateTransform      fusion_node.cpp:589-604
 *   dfb_prerotate_z90              prerotate(AngleAxisd(90deg, UnitZ))   fusion_node.cpp:774-777
 *   dfb_virtual_camera_transform   virtual camera pose                   fusion_node.cpp:168-180
 *   dfb_affine_to_float16          Pointcloud::setTransform cast         pointcloud.cpp:63-66
 *   dfb_fused_intrinsics           fused image intrinsics                fusion_node.cpp:155-165
 *   dfb_camera_info_to_intrinsics  cameraInfo2Intrinsics                 fusion_node.cpp:570-581
 *   dfb_depth_thresholds           NPP threshold values                  kernels.cu:357-359
 *   dfb_wire_unpack_depth/_color   DepthFrameset image payloads          camera_node.cpp:191-209, fusion_node.cpp:712-735
 *   dfb_wire_fused_image           fused sensor_msgs/Image header        fusion_node.cpp:799-808
 */
#include <cfloat>
#include <cmath>
#include <cstring>

#include "dfb/dfb.h"

namespace {

struct Quat {
	double x, y, z, w;
};

struct Mat3 {
	double a[3][3];
};

Mat3 linear_part(const double T[16]) {
	Mat3 m;
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) m.a[r][c] = T[4 * r + c];
	return m;
}

/* Eigen::Quaterniond(Matrix3d): quaternionbase_assign_impl<.,3,3>::run */
Quat quat_from(const Mat3& m) {
	double q[4]; /* x y z w */
	double t = m.a[0][0] + m.a[1][1] + m.a[2][2];
	if (t > 0.0) {
		t = std::sqrt(t + 1.0);
		q[3] = 0.5 * t;
		t = 0.5 / t;
		q[0] = (m.a[2][1] - m.a[1][2]) * t;
		q[1] = (m.a[0][2] - m.a[2][0]) * t;
		q[2] = (m.a[1][0] - m.a[0][1]) * t;
	} else {
		int i = 0;
		if (m.a[1][1] > m.a[0][0]) i = 1;
		if (m.a[2][2] > m.a[i][i]) i = 2;
		const int j = (i + 1) % 3, k = (j + 1) % 3;
		t = std::sqrt(m.a[i][i] - m.a[j][j] - m.a[k][k] + 1.0);
		q[i] = 0.5 * t;
		t = 0.5 / t;
		q[3] = (m.a[k][j] - m.a[j][k]) * t;
		q[j] = (m.a[j][i] + m.a[i][j]) * t;
		q[k] = (m.a[k][i] + m.a[i][k]) * t;
	}
	return Quat{q[0], q[1], q[2], q[3]};
}

/* Eigen::QuaternionBase::slerp */
Quat slerp(const Quat& a, const Quat& b, double t) {
	const double one = 1.0 - DBL_EPSILON;
	const double d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
	const double ad = std::fabs(d);
	double s0, s1;
	if (ad >= one) {
		s0 = 1.0 - t;
		s1 = t;
	} else {
		const double theta = std::acos(ad);
		const double st = std::sin(theta);
		s0 = std::sin((1.0 - t) * theta) / st;
		s1 = std::sin(t * theta) / st;
	}
	if (d < 0.0) s1 = -s1;
	return Quat{s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z, s0 * a.w + s1 * b.w};
}

/* Eigen::QuaternionBase::toRotationMatrix */
Mat3 to_matrix(const Quat& q) {
	const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
	const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
	const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
	const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
	Mat3 m;
	m.a[0][0] = 1.0 - (tyy + tzz);
	m.a[0][1] = txy - twz;
	m.a[0][2] = txz + twy;
	m.a[1][0] = txy + twz;
	m.a[1][1] = 1.0 - (txx + tzz);
	m.a[1][2] = tyz - twx;
	m.a[2][0] = txz - twy;
	m.a[2][1] = tyz + twx;
	m.a[2][2] = 1.0 - (txx + tyy);
	return m;
}

Mat3 inverse(const Mat3& m) {
	const double a = m.a[0][0], b = m.a[0][1], c = m.a[0][2];
	const double d = m.a[1][0], e = m.a[1][1], f = m.a[1][2];
	const double g = m.a[2][0], h = m.a[2][1], i = m.a[2][2];
	const double inv_det = 1.0 / (a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g));
	Mat3 r;
	r.a[0][0] = (e * i - f * h) * inv_det;
	r.a[0][1] = (c * h - b * i) * inv_det;
	r.a[0][2] = (b * f - c * e) * inv_det;
	r.a[1][0] = (f * g - d * i) * inv_det;
	r.a[1][1] = (a * i - c * g) * inv_det;
	r.a[1][2] = (c * d - a * f) * inv_det;
	r.a[2][0] = (d * h - e * g) * inv_det;
	r.a[2][1] = (b * g - a * h) * inv_det;
	r.a[2][2] = (a * e - b * d) * inv_det;
	return r;
}

Mat3 mul(const Mat3& x, const Mat3& y) {
	Mat3 r;
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) r.a[i][j] = x.a[i][0] * y.a[0][j] + x.a[i][1] * y.a[1][j] + x.a[i][2] * y.a[2][j];
	return r;
}

/* Eigen::AngleAxisd(angle, unit axis e_k).toRotationMatrix() */
Mat3 axis_rotation(int axis, double angle) {
	const double s = std::sin(angle), c = std::cos(angle);
	Mat3 m = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
	const int i = (axis + 1) % 3, j = (axis + 2) % 3;
	m.a[i][i] = c;
	m.a[j][j] = c;
	m.a[i][j] = -s;
	m.a[j][i] = s;
	m.a[axis][axis] = (1.0 - c) + c;
	return m;
}

void compose(const Mat3& R, const double t[3], double out[16]) {
	for (int r = 0; r < 3; ++r) {
		for (int c = 0; c < 3; ++c) out[4 * r + c] = R.a[r][c];
		out[4 * r + 3] = t[r];
	}
	out[12] = out[13] = out[14] = 0.0;
	out[15] = 1.0;
}

/* T <- [R 0; 0 1] * T  (Transform::prerotate acts on the whole affine block) */
void prerotate(double T[16], const Mat3& R) {
	for (int c = 0; c < 4; ++c) {
		const double v0 = T[c], v1 = T[4 + c], v2 = T[8 + c];
		for (int r = 0; r < 3; ++r) T[4 * r + c] = R.a[r][0] * v0 + R.a[r][1] * v1 + R.a[r][2] * v2;
	}
}

constexpr double kPi = 3.14159265358979323846;
double deg2rad(double d) { return d * kPi / 180.0; }

}  // namespace

extern "C" {

void dfb_interpolate_transform(const double left[16], const double right[16], double out[16]) {
	const Quat ql = quat_from(linear_part(left));
	const Quat qr = quat_from(linear_part(right));
	const Mat3 R = to_matrix(slerp(ql, qr, 0.5));
	double t[3];
	for (int i = 0; i < 3; ++i) t[i] = left[4 * i + 3] * 0.5 + right[4 * i + 3] * 0.5;
	if (qr.w < 0.0) { /* fusion_node.cpp:603: interpolated_transform.inverse() */
		const Mat3 Ri = inverse(R);
		double ti[3];
		for (int r = 0; r < 3; ++r) ti[r] = -(Ri.a[r][0] * t[0] + Ri.a[r][1] * t[1] + Ri.a[r][2] * t[2]);
		compose(Ri, ti, out);
	} else {
		compose(R, t, out);
	}
}

void dfb_prerotate_z90(double T[16]) { prerotate(T, axis_rotation(2, deg2rad(90.0))); }

void dfb_virtual_camera_transform(const double translation[3], const double rotation_deg[3], double out[16]) {
	/* fusion_node.cpp:170-179 */
	const Mat3 I = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
	const double zero[3] = {0, 0, 0};
	compose(I, zero, out);
	prerotate(out, axis_rotation(2, deg2rad(90.0)));
	const Mat3 R = mul(mul(axis_rotation(0, deg2rad(rotation_deg[0])), axis_rotation(1, deg2rad(rotation_deg[1]))),
	                   axis_rotation(2, deg2rad(rotation_deg[2])));
	for (int i = 0; i < 3; ++i) out[4 * i + 3] += -translation[i]; /* pretranslate(-translation) */
	prerotate(out, inverse(R));
}

void dfb_affine_to_float16(const double T[16], float out[16]) {
	for (int i = 0; i < 16; ++i) out[i] = static_cast<float>(T[i]);
}

void dfb_fused_intrinsics(const dfb_intrinsics* color_left, int vertical, dfb_intrinsics* out) {
	*out = *color_left;
	if (vertical) {
		out->width = color_left->height;
		out->height = color_left->width;
		out->fx = color_left->fy;
		out->fy = color_left->fx;
	}
	out->ppx = static_cast<float>(out->width / 2);
	out->ppy = static_cast<float>(out->height / 2);
}

void dfb_camera_info_to_intrinsics(uint32_t width, uint32_t height, const double K[9], const double D[5],
                                   dfb_intrinsics* out) {
	out->width = static_cast<int>(width);
	out->height = static_cast<int>(height);
	out->ppx = static_cast<float>(static_cast<int>(K[2]));
	out->ppy = static_cast<float>(static_cast<int>(K[5]));
	out->fx = static_cast<float>(static_cast<int>(K[0]));
	out->fy = static_cast<float>(static_cast<int>(K[4]));
	out->model = DFB_DISTORTION_BROWN_CONRADY;
	for (int i = 0; i < 5; ++i) out->coeffs[i] = static_cast<float>(D[i]);
}

void dfb_depth_thresholds(float min_depth, float max_depth, float depth_scale, uint16_t* lt, uint16_t* gt) {
	/* static_cast<Npp16u>(min_depth / depth_scale) in fp32, kernels.cu:357,359; clamped instead of UB */
	float a = min_depth / depth_scale;
	float b = max_depth / depth_scale;
	if (!(a > 0.0f)) a = 0.0f;
	if (a > 65535.0f) a = 65535.0f;
	if (!(b > 0.0f)) b = 0.0f;
	if (b > 65535.0f) b = 65535.0f;
	*lt = static_cast<uint16_t>(a);
	*gt = static_cast<uint16_t>(b);
}

}  // extern "C"

/* ---- wire formats ------------------------------------------------------------------------------------ */

namespace {
int check_image(const dfb_image_msg* msg, int width, int height, uint32_t bytes_per_pixel) {
	if (!msg || !msg->data || !msg->encoding || width <= 0 || height <= 0) return DFB_ERR_INVALID;
	if (msg->width != static_cast<uint32_t>(width) || msg->height != static_cast<uint32_t>(height)) return DFB_ERR_INVALID;
	if (msg->step < msg->width * bytes_per_pixel) return DFB_ERR_INVALID;
	if (msg->data_size < static_cast<size_t>(msg->step) * msg->height) return DFB_ERR_INVALID;
	return DFB_OK;
}
}  // namespace

int dfb_wire_unpack_depth(const dfb_image_msg* msg, int width, int height, uint16_t* dst) {
	const int rc = check_image(msg, width, height, 2);
	if (rc != DFB_OK || !dst) return DFB_ERR_INVALID;
	if (std::strcmp(msg->encoding, "mono16") != 0 && std::strcmp(msg->encoding, "16UC1") != 0) return DFB_ERR_INVALID;
	for (int v = 0; v < height; ++v) {
		const uint8_t* row = msg->data + static_cast<size_t>(v) * msg->step;
		uint16_t* out = dst + static_cast<size_t>(v) * width;
		if (!msg->is_bigendian) {
			std::memcpy(out, row, static_cast<size_t>(width) * 2); /* camera_node.cpp:196: is_bigendian = false */
		} else {
			for (int u = 0; u < width; ++u) out[u] = static_cast<uint16_t>((row[2 * u] << 8) | row[2 * u + 1]);
		}
	}
	return DFB_OK;
}

int dfb_wire_unpack_color(const dfb_image_msg* msg, int width, int height, uint8_t* dst) {
	const int rc = check_image(msg, width, height, 3);
	if (rc != DFB_OK || !dst) return DFB_ERR_INVALID;
	if (std::strcmp(msg->encoding, "rgb8") != 0) return DFB_ERR_INVALID; /* camera_node.cpp:207 */
	for (int v = 0; v < height; ++v)
		std::memcpy(dst + static_cast<size_t>(v) * width * 3, msg->data + static_cast<size_t>(v) * msg->step, static_cast<size_t>(width) * 3);
	return DFB_OK;
}

int dfb_wire_fused_image(const dfb_intrinsics* fused, const uint8_t* data, dfb_image_msg* out) {
	if (!fused || !data || !out || fused->width <= 0 || fused->height <= 0) return DFB_ERR_INVALID;
	out->width = static_cast<uint32_t>(fused->width);
	out->height = static_cast<uint32_t>(fused->height);
	out->is_bigendian = 0;
	out->step = out->width * 3;
	out->encoding = "rgb8";
	out->data = data;
	out->data_size = static_cast<size_t>(out->step) * out->height;
	return DFB_OK;
}
