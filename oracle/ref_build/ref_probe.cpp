// ref_probe — a pybind11 harness (our code) linked against the UNMODIFIED
// reference objects. It constructs a reference redmax::Simulation from an XML
// string and exposes the hot-path internals the reference keeps private or
// protected, so that golden vectors can be dumped and the CPU restatement in
// oracle/ can be validated against the reference itself:
//
//   BodySDFObj::query_SDF / query_dSDF / query_ddSDF / query_grid_location
//       (simulation/redmax/Body/BodySDFObj.cpp:123-205,379-383; private)
//   BodySDFObj::distance / collision (both overloads)   (.cpp:207-377; public)
//   collision_detection_general_SDF (CollisionDetection.cpp:124-147)
//   ForceGeneralSDFContact::computeForce(contacts, fm) and
//       computeForceWithDerivative(contacts, fm, Km, Dm)
//       (Force/ForceGeneralSDFContact.cpp:147-447; protected)
//   the public Force wrappers (.cpp:46-105) on caller-zeroed system arrays
//   Simulation::reset/forward/set_state/... for per-step trajectories.
//
// Access to private/protected members uses the `#define private public`
// trick in THIS translation unit only, after the std/Eigen/pybind headers and
// before the reference headers ("Oracle-access notes").
// No reference file is edited or copied. Test infrastructure only.

#include <pybind11/pybind11.h>
#include <pybind11/eigen.h>
#include <pybind11/stl.h>
#include <pybind11/numpy.h>

#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <Eigen/Sparse>
#include <vector>
#include <map>
#include <set>
#include <unordered_map>
#include <iostream>
#include <cmath>
#include <ctime>
#include <fstream>
#include <memory>
#include <chrono>
#include <sstream>
#include <iomanip>
#include <string>
#include <thread>

#define private public
#define protected public
#include "Simulation.h"
#include "Robot.h"
#include "Body/Body.h"
#include "Body/BodyMeshObj.h"
#include "Body/BodySDFObj.h"
#include "Joint/Joint.h"
#include "Force/Force.h"
#include "Force/ForceGeneralSDFContact.h"
#include "Force/ForceGroundContact.h"
#include "CollisionDetection/CollisionDetection.h"
#include "CollisionDetection/Contact.h"
#include "Mesh.h"
#undef private
#undef protected

namespace py = pybind11;
using namespace redmax;

namespace {

using RowMatX3 = Eigen::Matrix<double, Eigen::Dynamic, 3, Eigen::RowMajor>;
using RowMatX3i = Eigen::Matrix<int, Eigen::Dynamic, 3, Eigen::RowMajor>;

struct Probe {
    std::unique_ptr<Simulation> sim;

    Probe(const std::string& xml, const std::string& asset_folder) {
        try {
            sim.reset(new Simulation(xml, asset_folder, false));
        } catch (const std::string& msg) {
            throw std::runtime_error("reference threw: " + msg);
        } catch (const char* msg) {
            throw std::runtime_error(std::string("reference threw: ") + msg);
        }
    }

    Body* body(const std::string& name) {
        auto it = sim->_body_map.find(name);
        if (it == sim->_body_map.end()) throw std::runtime_error("no body " + name);
        return it->second;
    }
    BodySDFObj* sdf_body(const std::string& name) {
        BodySDFObj* b = dynamic_cast<BodySDFObj*>(body(name));
        if (!b) throw std::runtime_error("body " + name + " is not an SDF body");
        return b;
    }
    ForceGeneralSDFContact* sdf_force(int i) {
        if (i < 0 || i >= (int)sim->_robot->_forces.size()) throw std::runtime_error("bad force index");
        auto* f = dynamic_cast<ForceGeneralSDFContact*>(sim->_robot->_forces[i]);
        if (!f) throw std::runtime_error("force is not ForceGeneralSDFContact");
        return f;
    }

    int num_forces() { return (int)sim->_robot->_forces.size(); }

    std::vector<std::string> body_names() {
        std::vector<std::string> names;
        for (Body* b : sim->_robot->_bodies) names.push_back(b->_name);
        return names;
    }

    py::dict force_info(int i) {
        py::dict d;
        Force* f = sim->_robot->_forces.at(i);
        if (auto* c = dynamic_cast<ForceGeneralSDFContact*>(f)) {
            d["type"] = "general_SDF_contact";
            d["contact_body"] = c->_contact_body->_name;
            d["sdf_body"] = c->_SDF_body->_name;
            d["kn"] = c->_kn; d["kt"] = c->_kt; d["mu"] = c->_mu;
            d["damping"] = c->_damping; d["scale"] = c->_scale;
            d["contact_scale"] = c->_contact_body->_contact_scale;
            d["index_contact"] = c->_contact_body->_index[0];
            d["index_sdf"] = c->_SDF_body->_index[0];
        } else if (auto* g = dynamic_cast<ForceGroundContact*>(f)) {
            d["type"] = "ground_contact";
            d["contact_body"] = g->_contact_body->_name;
            d["kn"] = g->_kn; d["kt"] = g->_kt; d["mu"] = g->_mu; d["damping"] = g->_damping;
            d["E_g"] = Matrix4(g->_E_g);
            d["index_contact"] = g->_contact_body->_index[0];
        } else {
            d["type"] = "other";
        }
        return d;
    }

    py::dict body_info(const std::string& name) {
        Body* b = body(name);
        py::dict d;
        d["index"] = b->_index[0];
        d["mass"] = b->_mass;
        d["inertia"] = Vector6(b->_Inertia);
        d["contact_scale"] = b->_contact_scale;
        d["E_0i"] = Matrix4(b->_E_0i);
        d["phi"] = Vector6(b->_phi);
        d["E_ij"] = Matrix4(b->_E_ij);
        d["E_ji"] = Matrix4(b->_E_ji);
        d["f_external"] = Vector6(b->_f_external);
        d["joint_ndof"] = b->_joint->_ndof;
        return d;
    }

    RowMatX3 contact_points(const std::string& name) {
        std::vector<Vector3> pts = body(name)->get_contact_points();
        RowMatX3 out(pts.size(), 3);
        for (size_t i = 0; i < pts.size(); ++i) out.row(i) = pts[i].transpose();
        return out;
    }

    py::dict sdf_grid(const std::string& name) {
        BodySDFObj* b = sdf_body(name);
        py::dict d;
        d["ni"] = b->_SDF.ni; d["nj"] = b->_SDF.nj; d["nk"] = b->_SDF.nk;
        d["min_box"] = Vector3(b->_min_box); d["max_box"] = Vector3(b->_max_box);
        d["dx"] = b->_dx; d["dy"] = b->_dy; d["dz"] = b->_dz;
        d["col_th"] = b->_col_th;
        // linear layout exactly as stored: a[i + ni*(j + nj*k)] (sdfgen/array3.h:70-80)
        py::array_t<double> grid(b->_SDF.a.size());
        std::copy(b->_SDF.a.begin(), b->_SDF.a.end(), grid.mutable_data());
        d["grid"] = grid;
        return d;
    }

    void set_body_pose(const std::string& name, const Matrix4& E, const Vector6& phi) {
        Body* b = body(name);
        b->_E_0i = E;
        b->_E_i0 = math::Einv(E);
        b->_phi = phi;
    }
    // Overwrite the hot-path inputs of a body (public data members) so that the
    // reference evaluates exactly the arrays the CUDA path is given.
    void set_contact_points(const std::string& name, const RowMatX3& X) {
        Body* b = body(name);
        b->_contact_points.resize(X.rows());
        for (int i = 0; i < X.rows(); ++i) b->_contact_points[i] = Vector3(X.row(i).transpose());
    }
    void set_sdf_grid(const std::string& name, int ni, int nj, int nk, const Vector3& mn, const Vector3& mx,
                      double dx, double dy, double dz, py::array_t<double, py::array::c_style | py::array::forcecast> grid) {
        BodySDFObj* b = sdf_body(name);
        if ((size_t)grid.size() != (size_t)ni * nj * nk) throw std::runtime_error("grid size mismatch");
        b->_SDF.resize(ni, nj, nk);
        std::copy(grid.data(), grid.data() + grid.size(), b->_SDF.a.begin());
        b->_min_box = mn; b->_max_box = mx; b->_dx = dx; b->_dy = dy; b->_dz = dz;
    }
    void set_col_th(const std::string& name, double col_th) { sdf_body(name)->_col_th = col_th; }
    // Overwrite the mass properties / body frame the reference derived from the mesh
    // (BodyMeshObj.cpp:70-177 -> Body.cpp:63-66,88-91) so that both arms integrate
    // exactly the same rigid bodies.
    void set_body_frame(const std::string& name, const Matrix4& E_ji, const Vector6& inertia, double mass) {
        Body* b = body(name);
        b->_E_ji = E_ji;
        b->_E_ij = math::Einv(E_ji);
        b->_A_ji = math::Ad(b->_E_ji);
        b->_A_ij = math::Ad(b->_E_ij);
        b->_Inertia = inertia;
        b->_mass = mass;
    }
    py::dict joint_info(const std::string& body_name) {
        Joint* j = body(body_name)->_joint;
        py::dict d;
        d["ndof"] = j->_ndof;
        d["E_pj_0"] = Matrix4(j->_E_pj_0);
        d["index"] = j->_ndof > 0 ? j->_index[0] : -1;
        return d;
    }
    // Wall time of Simulation::forward(n) (the stepping loop the planners drive).
    double time_forward(int n) {
        auto t0 = std::chrono::steady_clock::now();
        sim->forward(n, false, false);
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    }
    void set_force_params(int i, double kn, double kt, double mu, double damping, double scale) {
        if (auto* g = dynamic_cast<ForceGroundContact*>(sim->_robot->_forces.at(i))) {
            g->_kn = kn; g->_kt = kt; g->_mu = mu; g->_damping = damping;  // no _scale on the ground force
            return;
        }
        auto* f = sdf_force(i);
        f->_kn = kn; f->_kt = kt; f->_mu = mu; f->_damping = damping; f->_scale = scale;
    }
    void set_contact_scale(const std::string& name, double s) { body(name)->_contact_scale = s; }

    // ---- SDF sampling (private members of BodySDFObj) ----
    Eigen::VectorXd query_sdf(const std::string& name, const RowMatX3& X, bool outside_accurate) {
        BodySDFObj* b = sdf_body(name);
        Eigen::VectorXd out(X.rows());
        for (int i = 0; i < X.rows(); ++i) out[i] = b->query_SDF(Vector3(X.row(i).transpose()), outside_accurate);
        return out;
    }
    RowMatX3 query_dsdf(const std::string& name, const RowMatX3& X) {
        BodySDFObj* b = sdf_body(name);
        RowMatX3 out(X.rows(), 3);
        for (int i = 0; i < X.rows(); ++i) out.row(i) = b->query_dSDF(Vector3(X.row(i).transpose())).transpose();
        return out;
    }
    py::array_t<double> query_ddsdf(const std::string& name, const RowMatX3& X) {
        BodySDFObj* b = sdf_body(name);
        py::array_t<double> out({(py::ssize_t)X.rows(), (py::ssize_t)3, (py::ssize_t)3});
        auto o = out.mutable_unchecked<3>();
        for (int i = 0; i < X.rows(); ++i) {
            Matrix3 H = b->query_ddSDF(Vector3(X.row(i).transpose()));
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o(i, r, c) = H(r, c);
        }
        return out;
    }
    RowMatX3i grid_location(const std::string& name, const RowMatX3& X) {
        BodySDFObj* b = sdf_body(name);
        RowMatX3i out(X.rows(), 3);
        for (int i = 0; i < X.rows(); ++i) out.row(i) = b->query_grid_location(Vector3(X.row(i).transpose())).transpose();
        return out;
    }
    Eigen::VectorXd distance(const std::string& name, const RowMatX3& Xw) {
        BodySDFObj* b = sdf_body(name);
        Eigen::VectorXd out(Xw.rows());
        for (int i = 0; i < Xw.rows(); ++i) out[i] = b->distance(Vector3(Xw.row(i).transpose()));
        return out;
    }

    // ---- detection: every point, exactly as collision_detection_general_SDF walks them ----
    py::dict detect(int fi) {
        auto* f = sdf_force(fi);
        Body* A = f->_contact_body;
        BodySDFObj* B = f->_SDF_body;
        std::vector<Vector3> pts = A->get_contact_points();
        const int n = (int)pts.size();
        Eigen::VectorXd d(n);
        RowMatX3 xw(n, 3), xl(n, 3);
        RowMatX3i cell(n, 3);
        Matrix4 E_w1 = A->_E_0i;
        Matrix3 R2 = B->_E_0i.topLeftCorner(3, 3);
        Vector3 p2 = B->_E_0i.topRightCorner(3, 1);
        for (int i = 0; i < n; ++i) {
            Vector3 xw1 = E_w1.topLeftCorner(3, 3) * pts[i] + E_w1.topRightCorner(3, 1);
            d[i] = B->distance(xw1);
            Vector3 x = R2.transpose() * (xw1 - p2);
            xw.row(i) = xw1.transpose();
            xl.row(i) = x.transpose();
            cell.row(i) = B->query_grid_location(x).transpose();
        }
        // the reference's own compacted list (also records contact history)
        std::vector<Contact> contacts;
        collision_detection_general_SDF(A, B, contacts);
        std::vector<int> ids; ids.reserve(contacts.size());
        for (auto& c : contacts) ids.push_back(c._id);
        py::dict out;
        out["d"] = d; out["xw"] = xw; out["x_local"] = xl; out["cell"] = cell; out["active_ids"] = ids;
        return out;
    }

    // ground contact: the reference's own contact list (ids, d) for force fi
    py::dict ground_detect(int fi) {
        auto* g = dynamic_cast<ForceGroundContact*>(sim->_robot->_forces.at(fi));
        if (!g) throw std::runtime_error("force is not ForceGroundContact");
        std::vector<Contact> contacts;
        collision_detection_ground_object(g->_E_g, g->_contact_body, contacts);
        std::vector<int> ids; std::vector<double> d;
        for (auto& c : contacts) { ids.push_back(c._id); d.push_back(c._d); }
        py::dict out;
        out["active_ids"] = ids; out["d"] = d;
        return out;
    }

    // ---- protected per-pair kernels on the reference's own contact list ----
    Eigen::VectorXd force_sub(int fi) {
        auto* f = sdf_force(fi);
        std::vector<Contact> contacts;
        collision_detection_general_SDF(f->_contact_body, f->_SDF_body, contacts);
        VectorX fm;
        f->computeForce(contacts, fm, false);
        return fm;
    }
    py::tuple force_deriv_sub(int fi) {
        auto* f = sdf_force(fi);
        std::vector<Contact> contacts;
        collision_detection_general_SDF(f->_contact_body, f->_SDF_body, contacts);
        VectorX fm; MatrixX Km, Dm;
        f->computeForceWithDerivative(contacts, fm, Km, Dm, false);
        return py::make_tuple(fm, Km, Dm);
    }
    // per-contact terms: evaluate the protected kernels on single-contact lists
    py::tuple force_deriv_per_contact(int fi) {
        auto* f = sdf_force(fi);
        std::vector<Contact> contacts;
        collision_detection_general_SDF(f->_contact_body, f->_SDF_body, contacts);
        const int n = (int)contacts.size();
        py::array_t<double> FM({(py::ssize_t)n, (py::ssize_t)12});
        py::array_t<double> KM({(py::ssize_t)n, (py::ssize_t)12, (py::ssize_t)12});
        py::array_t<double> DM({(py::ssize_t)n, (py::ssize_t)12, (py::ssize_t)12});
        auto fmv = FM.mutable_unchecked<2>();
        auto kmv = KM.mutable_unchecked<3>();
        auto dmv = DM.mutable_unchecked<3>();
        std::vector<int> ids;
        for (int c = 0; c < n; ++c) {
            std::vector<Contact> one(1, contacts[c]);
            VectorX fm; MatrixX Km, Dm;
            f->computeForceWithDerivative(one, fm, Km, Dm, false);
            for (int r = 0; r < 12; ++r) {
                fmv(c, r) = fm[r];
                for (int k = 0; k < 12; ++k) { kmv(c, r, k) = Km(r, k); dmv(c, r, k) = Dm(r, k); }
            }
            ids.push_back(contacts[c]._id);
        }
        return py::make_tuple(ids, FM, KM, DM);
    }

    // ---- public Force wrappers on zeroed system arrays (accumulate semantics) ----
    Eigen::VectorXd force_public(int fi) {
        VectorX fm = VectorX::Zero(sim->_ndof_m), fr = VectorX::Zero(sim->_ndof_r);
        sim->_robot->_forces.at(fi)->computeForce(fm, fr, false);
        return fm;
    }
    py::tuple force_deriv_public(int fi) {
        const int nm = sim->_ndof_m, nr = sim->_ndof_r;
        VectorX fm = VectorX::Zero(nm), fr = VectorX::Zero(nr);
        MatrixX Km = MatrixX::Zero(nm, nm), Dm = MatrixX::Zero(nm, nm);
        MatrixX Kr = MatrixX::Zero(nr, nr), Dr = MatrixX::Zero(nr, nr);
        sim->_robot->_forces.at(fi)->computeForceWithDerivative(fm, fr, Km, Dm, Kr, Dr, false);
        return py::make_tuple(fm, Km, Dm);
    }

    // ---- BodySDFObj::collision (public) ----
    py::dict collision(const std::string& name, const Vector3& xw, const Vector3& xw_dot, bool deriv) {
        BodySDFObj* b = sdf_body(name);
        dtype d, ddot; Vector3 n, tdot, xi2;
        py::dict out;
        if (!deriv) {
            b->collision(xw, xw_dot, d, n, ddot, tdot, xi2);
        } else {
            RowVector3 dd_dxw, dddot_dxw, dddot_dxwdot;
            RowVector6 dd_dq2, dddot_dq2, dddot_dphi2;
            Matrix3 dn_dxw, dtdot_dxw, dtdot_dxwdot, dxi2_dxw;
            Matrix36 dn_dq2, dtdot_dq2, dtdot_dphi2, dxi2_dq2;
            b->collision(xw, xw_dot, d, n, ddot, tdot, xi2, dd_dxw, dd_dq2, dn_dxw, dn_dq2,
                         dddot_dxw, dddot_dxwdot, dddot_dq2, dddot_dphi2,
                         dtdot_dxw, dtdot_dxwdot, dtdot_dq2, dtdot_dphi2, dxi2_dxw, dxi2_dq2);
            out["dd_dxw"] = dd_dxw; out["dd_dq2"] = dd_dq2; out["dn_dxw"] = dn_dxw; out["dn_dq2"] = dn_dq2;
            out["dddot_dxw"] = dddot_dxw; out["dddot_dxwdot"] = dddot_dxwdot;
            out["dddot_dq2"] = dddot_dq2; out["dddot_dphi2"] = dddot_dphi2;
            out["dtdot_dxw"] = dtdot_dxw; out["dtdot_dxwdot"] = dtdot_dxwdot;
            out["dtdot_dq2"] = dtdot_dq2; out["dtdot_dphi2"] = dtdot_dphi2;
            out["dxi2_dxw"] = dxi2_dxw; out["dxi2_dq2"] = dxi2_dq2;
        }
        out["d"] = d; out["n"] = n; out["ddot"] = ddot; out["tdot"] = tdot; out["xi2"] = xi2;
        return out;
    }

    // ---- whole-system quantities (Robot level) ----
    py::tuple robot_force_deriv() {
        const int nm = sim->_ndof_m, nr = sim->_ndof_r;
        VectorX fm = VectorX::Zero(nm), fr = VectorX::Zero(nr);
        MatrixX Km = MatrixX::Zero(nm, nm), Dm = MatrixX::Zero(nm, nm);
        MatrixX Kr = MatrixX::Zero(nr, nr), Dr = MatrixX::Zero(nr, nr);
        sim->_robot->computeForceWithDerivative(fm, fr, Km, Dm, Kr, Dr, false);
        return py::make_tuple(fm, fr, Km, Dm, Kr, Dr);
    }
    py::tuple robot_force() {
        VectorX fm = VectorX::Zero(sim->_ndof_m), fr = VectorX::Zero(sim->_ndof_r);
        sim->_robot->computeForce(fm, fr);
        return py::make_tuple(fm, fr);
    }
    py::tuple compute_matrices(bool deriv) {
        MatrixX M; VectorX f;
        if (!deriv) { sim->computeMatrices(M, f); return py::make_tuple(M, f); }
        JacobianMatrixVector dM_dq; MatrixX K, D;
        sim->computeMatrices(M, f, dM_dq, K, D);
        return py::make_tuple(M, f, K, D);
    }
    py::tuple joint_jacobian() {
        MatrixX J, Jdot;
        sim->_robot->computeJointJacobian(J, Jdot);
        return py::make_tuple(J, Jdot);
    }
    Eigen::VectorXd maximal_mass() {
        VectorX Mm; sim->_robot->computeMaximalMassMatrix(Mm); return Mm;
    }

    // ---- simulation driving ----
    void reset() { sim->reset(false, false); }
    void forward(int n) { sim->forward(n, false, false); }
    Eigen::VectorXd get_q() { return sim->get_q(); }
    Eigen::VectorXd get_qdot() { return sim->get_qdot(); }
    void set_state(const Eigen::VectorXd& q, const Eigen::VectorXd& qdot) { sim->set_state(q, qdot); }
    void set_state_init(const Eigen::VectorXd& q, const Eigen::VectorXd& qdot) { sim->set_state_init(q, qdot); }
    void update_robot() { sim->update_robot(false); }
    bool is_converged() { return sim->is_converged(); }
    int ndof_r() { return sim->_ndof_r; }
    int ndof_m() { return sim->_ndof_m; }
    double timestep() { return sim->_options->_h; }
    Eigen::Vector3d gravity() { return sim->_options->_gravity; }
    void set_body_external_force(const std::string& name, const Eigen::VectorXd& f) { sim->set_body_external_force(name, f); }
    std::vector<std::string> get_contact_bodies(const std::string& name) { return sim->get_contact_bodies(name); }
    void clear_all_contact_bodies() { sim->clear_all_contact_bodies(); }
    Eigen::VectorXd get_joint_q(const std::string& name) { return sim->get_joint_q(name); }
    void set_joint_q(const std::string& name, const Eigen::VectorXd& q) { sim->set_joint_q(name, q); }
    Eigen::VectorXd get_joint_qdot(const std::string& name) { return sim->get_joint_qdot(name); }
    void zero_joint_qdot(const std::string& name) { sim->zero_joint_qdot(name); }
    Matrix4 get_body_E0i(const std::string& name) { return sim->get_body_E0i(name); }
    Matrix3X get_body_vertices(const std::string& name, bool world_frame) { return sim->get_body_vertices(name, world_frame); }
    Eigen::Matrix<int, 3, Eigen::Dynamic> get_body_faces(const std::string& name) { return sim->get_body_faces(name); }

    void reset_time_report() { sim->_robot->reset_time_report(); }

    // Time `reps` evaluations of the public wrappers (the hot path exactly as
    // Robot::computeForce[WithDerivative] drives it). Returns seconds.
    double time_force_evals(int reps, bool deriv) {
        const int nm = sim->_ndof_m, nr = sim->_ndof_r;
        auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r) {
            VectorX fm = VectorX::Zero(nm), fr = VectorX::Zero(nr);
            if (deriv) {
                MatrixX Km = MatrixX::Zero(nm, nm), Dm = MatrixX::Zero(nm, nm);
                MatrixX Kr = MatrixX::Zero(nr, nr), Dr = MatrixX::Zero(nr, nr);
                for (Force* f : sim->_robot->_forces) f->computeForceWithDerivative(fm, fr, Km, Dm, Kr, Dr, false);
            } else {
                for (Force* f : sim->_robot->_forces) f->computeForce(fm, fr, false);
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    }
};

}  // namespace

// sdfgen::make_level_set3 of the reference, called directly (externals/sdfgen/makelevelset3.cpp)
static py::array_t<float> ref_make_level_set3(const Eigen::Matrix<unsigned, Eigen::Dynamic, 3, Eigen::RowMajor>& tri,
                                              const Eigen::Matrix<float, Eigen::Dynamic, 3, Eigen::RowMajor>& x,
                                              const Eigen::Vector3f& origin, float dx, float dy, float dz, int ni, int nj,
                                              int nk, int exact_band) {
    std::vector<sdfgen::Vec3ui> faces;
    std::vector<sdfgen::Vec3f> verts;
    for (int i = 0; i < tri.rows(); ++i) faces.push_back(sdfgen::Vec3ui(tri(i, 0), tri(i, 1), tri(i, 2)));
    for (int i = 0; i < x.rows(); ++i) verts.push_back(sdfgen::Vec3f(x(i, 0), x(i, 1), x(i, 2)));
    sdfgen::Array3f phi;
    sdfgen::make_level_set3(faces, verts, sdfgen::Vec3f(origin[0], origin[1], origin[2]), dx, dy, dz, ni, nj, nk, phi, exact_band);
    py::array_t<float> out((size_t)ni * nj * nk);
    std::memcpy(out.mutable_data(), &phi.a[0], sizeof(float) * (size_t)ni * nj * nk);
    return out;
}

PYBIND11_MODULE(ref_probe, m) {
    m.doc() = "probe harness over the unmodified reference RedMax objects (test infrastructure)";
    m.def("make_level_set3", &ref_make_level_set3, py::arg("tri"), py::arg("x"), py::arg("origin"), py::arg("dx"),
          py::arg("dy"), py::arg("dz"), py::arg("ni"), py::arg("nj"), py::arg("nk"), py::arg("exact_band") = 1);
    py::class_<Probe>(m, "Probe")
        .def(py::init<const std::string&, const std::string&>(), py::arg("xml"), py::arg("asset_folder"))
        .def("num_forces", &Probe::num_forces)
        .def("body_names", &Probe::body_names)
        .def("force_info", &Probe::force_info)
        .def("body_info", &Probe::body_info)
        .def("contact_points", &Probe::contact_points)
        .def("sdf_grid", &Probe::sdf_grid)
        .def("set_body_pose", &Probe::set_body_pose)
        .def("set_contact_points", &Probe::set_contact_points)
        .def("set_sdf_grid", &Probe::set_sdf_grid)
        .def("set_col_th", &Probe::set_col_th)
        .def("set_body_frame", &Probe::set_body_frame)
        .def("joint_info", &Probe::joint_info)
        .def("time_forward", &Probe::time_forward, py::call_guard<py::gil_scoped_release>())
        .def("set_force_params", &Probe::set_force_params)
        .def("set_contact_scale", &Probe::set_contact_scale)
        .def("query_sdf", &Probe::query_sdf, py::arg("name"), py::arg("X"), py::arg("outside_accurate") = false)
        .def("query_dsdf", &Probe::query_dsdf)
        .def("query_ddsdf", &Probe::query_ddsdf)
        .def("grid_location", &Probe::grid_location)
        .def("distance", &Probe::distance)
        .def("detect", &Probe::detect)
        .def("ground_detect", &Probe::ground_detect)
        .def("force_sub", &Probe::force_sub)
        .def("force_deriv_sub", &Probe::force_deriv_sub)
        .def("force_deriv_per_contact", &Probe::force_deriv_per_contact)
        .def("force_public", &Probe::force_public)
        .def("force_deriv_public", &Probe::force_deriv_public)
        .def("collision", &Probe::collision, py::arg("name"), py::arg("xw"), py::arg("xw_dot"), py::arg("deriv") = false)
        .def("robot_force", &Probe::robot_force)
        .def("robot_force_deriv", &Probe::robot_force_deriv)
        .def("compute_matrices", &Probe::compute_matrices)
        .def("joint_jacobian", &Probe::joint_jacobian)
        .def("maximal_mass", &Probe::maximal_mass)
        .def("reset", &Probe::reset)
        .def("forward", &Probe::forward)
        .def("get_q", &Probe::get_q)
        .def("get_qdot", &Probe::get_qdot)
        .def("set_state", &Probe::set_state)
        .def("set_state_init", &Probe::set_state_init)
        .def("update_robot", &Probe::update_robot)
        .def("is_converged", &Probe::is_converged)
        .def("ndof_r", &Probe::ndof_r)
        .def("ndof_m", &Probe::ndof_m)
        .def("timestep", &Probe::timestep)
        .def("gravity", &Probe::gravity)
        .def("set_body_external_force", &Probe::set_body_external_force)
        .def("get_contact_bodies", &Probe::get_contact_bodies)
        .def("clear_all_contact_bodies", &Probe::clear_all_contact_bodies)
        .def("get_joint_q", &Probe::get_joint_q)
        .def("set_joint_q", &Probe::set_joint_q)
        .def("get_joint_qdot", &Probe::get_joint_qdot)
        .def("zero_joint_qdot", &Probe::zero_joint_qdot)
        .def("get_body_E0i", &Probe::get_body_E0i)
        .def("get_body_vertices", &Probe::get_body_vertices, py::arg("name"), py::arg("world_frame") = false)
        .def("get_body_faces", &Probe::get_body_faces)
        .def("reset_time_report", &Probe::reset_time_report)
        .def("time_force_evals", &Probe::time_force_evals, py::arg("reps"), py::arg("deriv"),
             py::call_guard<py::gil_scoped_release>());
}
