// A small, functional, one-process stand-in for the slice of DOLFIN 2019.1.0 that the reference's
// src/fHSL.{h,cpp}, src/Expressions.h and the FFC-generated form headers (fenics/*.h) use, written here so that
// the reference's own class `fenicsInterface` can be compiled IN PLACE from /root/reference and run as a parity
// pin for the oracle's P1 path (oracle/Makefile -> oracle/_ref/libeq_fenics_ref.so).  TEST INFRASTRUCTURE ONLY.
//
// What runs unmodified on top of it: the reference's boundary-condition decoding (createHSL), Robin rates, form
// parameter wiring, step orchestration (trap solve -> wall flux -> channel sub-steps -> flux functional), and the
// reference's own FFC-generated element kernels (tabulate_tensor), which this shim's assembler calls.
// What is restated here from DOLFIN's documented behaviour [ext]: RectangleMesh("right") / IntervalMesh
// vertex and cell numbering, P1 dof numbering (identity to vertices on one process), coefficient restriction
// (vertex values for P1, the value for Real), facet marking by SubDomain::inside on all facet vertices and the
// midpoint, DirichletBC::apply on the assembled system (identity rows, columns kept, topological search over
// boundary facets), assembly over cells and marked exterior facets, and the linear solve (banded LU, no
// pivoting needed for these diagonally dominant systems; DOLFIN's default is a sparse direct LU).
#ifndef EQ_B200_SHIM_DOLFIN_MINI_H
#define EQ_B200_SHIM_DOLFIN_MINI_H
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <ufc.h>
#include "mpi.h"

namespace dolfin
{
const double DOLFIN_EPS = 3.0e-16;
const double DOLFIN_PI = 3.14159265358979323846;
enum { DBG = 10, TRACE = 13, PROGRESS = 16, INFO = 20, WARNING = 30, ERROR = 40, CRITICAL = 50 };
inline void set_log_level(int) {}
inline void dolfin_error(std::string where, std::string what, std::string why, ...)
{
  throw std::runtime_error("dolfin shim: " + where + ": " + what + ": " + why);
}
// DOLFIN: near(x, x0, eps) == between(x, x0 - eps, x0 + eps)
inline bool near(double x, double x0, double eps = DOLFIN_EPS) { return x0 - eps <= x && x <= x0 + eps; }

template <typename T>
class Array
{
public:
  explicit Array(std::size_t n) : _n(n), _own(new T[n]()), _x(_own.get()) {}
  Array(std::size_t n, T* x) : _n(n), _x(x) {}
  std::size_t size() const { return _n; }
  const T& operator[](std::size_t i) const { return _x[i]; }
  T& operator[](std::size_t i) { return _x[i]; }
  T* data() { return _x; }
  const T* data() const { return _x; }
private:
  std::size_t _n;
  std::unique_ptr<T[]> _own;
  T* _x;
};

class Point
{
public:
  Point(double x = 0.0, double y = 0.0, double z = 0.0) : c{x, y, z} {}
  double x() const { return c[0]; }
  double y() const { return c[1]; }
  double operator[](std::size_t i) const { return c[i]; }
  double c[3];
};

// ---- parameters["ghost_mode"] -------------------------------------------------------------------------------
class ParameterValue
{
public:
  ParameterValue& operator=(const std::string& s) { v = s; return *this; }
  ParameterValue& operator=(const char* s) { v = s; return *this; }
  operator std::string() const { return v; }
  std::string v;
};
class Parameters
{
public:
  ParameterValue& operator[](const std::string& k) { return m[k]; }
  std::map<std::string, ParameterValue> m;
};
extern Parameters parameters;
template <class... A> inline void info(A...) {}
inline std::string dolfin_version() { return "2019.1.0 (interface shim)"; }
inline std::string ufc_signature() { return "shim"; }
inline std::string git_commit_hash() { return "shim"; }
class SubSystemsManager
{
public:
  static SubSystemsManager& singleton() { static SubSystemsManager s; return s; }
  static void finalize() {}
};

// ---- meshes ---------------------------------------------------------------------------------------------------
class MeshTopology
{
public:
  explicit MeshTopology(std::size_t d = 0) : _d(d) {}
  std::size_t dim() const { return _d; }
private:
  std::size_t _d;
};

// A facet of the boundary: the cell it belongs to, its local number in that cell (UFC triangle: facet f lies opposite
// local vertex f; UFC interval: facet f is local vertex f), its vertices.
struct BoundaryFacet { std::size_t cell, local; std::vector<std::size_t> v; };

class Mesh
{
public:
  Mesh() {}
  virtual ~Mesh() {}
  std::size_t num_vertices() const { return _x.size() / _gdim; }
  std::size_t num_cells() const { return _cells.size() / (_tdim + 1); }
  std::vector<double>& coordinates() { return _x; }
  const std::vector<double>& coordinates() const { return _x; }
  const MeshTopology& topology() const { return _top; }
  std::size_t gdim() const { return _gdim; }
  std::size_t tdim() const { return _tdim; }
  const std::size_t* cell(std::size_t c) const { return &_cells[c * (_tdim + 1)]; }
  const std::vector<BoundaryFacet>& boundary_facets() const { return _bfacets; }
  // global facet index of a boundary facet = its position in boundary_facets() (interior facets are never
  // marked by the reference)
protected:
  std::size_t _gdim = 0, _tdim = 0;
  MeshTopology _top;
  std::vector<double> _x;          // vertex coordinates, vertex-major
  std::vector<std::size_t> _cells; // tdim+1 vertices per cell, ascending global index (UFC ordering)
  std::vector<BoundaryFacet> _bfacets;
};

// RectangleMesh(comm, p0, p1, nx, ny, "right"): vertices row-major v = iy*(nx+1) + ix; every square
// (v0 = BL, v1 = BR, v2 = TL, v3 = TR) is split by the BL->TR diagonal into (v0, v1, v3) and (v0, v2, v3).
class RectangleMesh : public Mesh
{
public:
  RectangleMesh(MPI_Comm, const Point& p0, const Point& p1, std::size_t nx, std::size_t ny, std::string diagonal = "right")
  {
    if (diagonal != "right") dolfin_error("RectangleMesh", "create mesh", "the shim builds the \"right\" diagonal only");
    _gdim = 2; _tdim = 2; _top = MeshTopology(2);
    const double a = p0.x(), b = p1.x(), c = p0.y(), d = p1.y();
    for (std::size_t iy = 0; iy <= ny; ++iy)
    {
      const double y = c + ((static_cast<double>(iy)) * (d - c) / static_cast<double>(ny));
      for (std::size_t ix = 0; ix <= nx; ++ix)
      {
        const double x = a + ((static_cast<double>(ix)) * (b - a) / static_cast<double>(nx));
        _x.push_back(x); _x.push_back(y);
      }
    }
    for (std::size_t iy = 0; iy < ny; ++iy)
      for (std::size_t ix = 0; ix < nx; ++ix)
      {
        const std::size_t v0 = iy * (nx + 1) + ix, v1 = v0 + 1, v2 = v0 + (nx + 1), v3 = v1 + (nx + 1);
        _cells.insert(_cells.end(), {v0, v1, v3});
        _cells.insert(_cells.end(), {v0, v2, v3});
      }
    // boundary facets: an edge is on the boundary iff it belongs to exactly one cell
    std::map<std::pair<std::size_t, std::size_t>, std::vector<std::pair<std::size_t, std::size_t>>> edges;
    for (std::size_t cidx = 0; cidx < num_cells(); ++cidx)
    {
      const std::size_t* v = cell(cidx);
      for (std::size_t f = 0; f < 3; ++f)
      {
        std::size_t e0 = v[(f + 1) % 3], e1 = v[(f + 2) % 3];
        if (e0 > e1) std::swap(e0, e1);
        edges[{e0, e1}].push_back({cidx, f});
      }
    }
    for (auto& e : edges)
      if (e.second.size() == 1)
        _bfacets.push_back(BoundaryFacet{e.second[0].first, e.second[0].second, {e.first.first, e.first.second}});
  }
};

class IntervalMesh : public Mesh
{
public:
  IntervalMesh(MPI_Comm, std::size_t n, double a, double b)
  {
    _gdim = 1; _tdim = 1; _top = MeshTopology(1);
    for (std::size_t i = 0; i <= n; ++i) _x.push_back(a + (static_cast<double>(i)) * (b - a) / static_cast<double>(n));
    for (std::size_t i = 0; i < n; ++i) _cells.insert(_cells.end(), {i, i + 1});
    // the facets of an interval are its vertices and are numbered like them (UFC/FIAT reference interval: facet 0 at
    // X = 0 with normal -1, facet 1 at X = 1 with normal +1; the generated PI0[facet] tables select local dof `facet`)
    _bfacets.push_back(BoundaryFacet{0, 0, {0}});
    _bfacets.push_back(BoundaryFacet{n - 1, 1, {n}});
  }
};

template <typename T>
class MeshFunction
{
public:
  MeshFunction(std::shared_ptr<const Mesh> mesh, std::size_t dim, const T& value)
    : _mesh(mesh), _dim(dim), _values(mesh->boundary_facets().size(), value) {}
  std::shared_ptr<const Mesh> mesh() const { return _mesh; }
  std::size_t dim() const { return _dim; }
  // boundary facets only (see Mesh::boundary_facets)
  T& boundary_value(std::size_t k) { return _values[k]; }
  const T& boundary_value(std::size_t k) const { return _values[k]; }
private:
  std::shared_ptr<const Mesh> _mesh;
  std::size_t _dim;
  std::vector<T> _values;
};

class SubDomain
{
public:
  virtual ~SubDomain() {}
  virtual bool inside(const Array<double>& x, bool on_boundary) const { return false; }
  // DOLFIN marks an entity when all of its vertices and its midpoint are inside
  bool facet_inside(const Mesh& mesh, const BoundaryFacet& f) const
  {
    const std::size_t g = mesh.gdim();
    std::vector<double> mid(g, 0.0);
    for (std::size_t v : f.v)
    {
      std::vector<double> p(mesh.coordinates().begin() + v * g, mesh.coordinates().begin() + (v + 1) * g);
      Array<double> x(g, p.data());
      if (!inside(x, true)) return false;
      for (std::size_t k = 0; k < g; ++k) mid[k] += p[k] / static_cast<double>(f.v.size());
    }
    Array<double> xm(g, mid.data());
    return inside(xm, true);
  }
  void mark(MeshFunction<std::size_t>& mf, std::size_t value) const
  {
    const Mesh& mesh = *mf.mesh();
    for (std::size_t k = 0; k < mesh.boundary_facets().size(); ++k)
      if (facet_inside(mesh, mesh.boundary_facets()[k])) mf.boundary_value(k) = value;
  }
};

// ---- functions ------------------------------------------------------------------------------------------------
class GenericFunction
{
public:
  virtual ~GenericFunction() {}
  virtual void eval(Array<double>& values, const Array<double>& x) const = 0;
  virtual std::size_t value_size() const { return 1; }
  double at(const double* x, std::size_t gdim) const
  {
    double v[3] = {0.0, 0.0, 0.0};
    Array<double> values(value_size() < 3 ? 3 : value_size(), v);
    double p[3] = {0.0, 0.0, 0.0};
    for (std::size_t k = 0; k < gdim; ++k) p[k] = x[k];
    Array<double> xx(gdim, p);
    eval(values, xx);
    return v[0];
  }
};

class Expression : public GenericFunction
{
public:
  Expression() : _vs(1) {}
  explicit Expression(std::size_t dim) : _vs(dim) {}
  void eval(Array<double>&, const Array<double>&) const override {}
  std::size_t value_size() const override { return _vs; }
private:
  std::size_t _vs;
};

class Constant : public Expression
{
public:
  explicit Constant(double v) : _v(v) {}
  void eval(Array<double>& values, const Array<double>&) const override { values[0] = _v; }
  operator double() const { return _v; }
  const Constant& operator=(double v) { _v = v; return *this; }
private:
  double _v;
};

class GenericVector
{
public:
  explicit GenericVector(std::size_t n = 0) : v(n, 0.0) {}
  void set_local(const std::vector<double>& x) { v = x; }
  void get_local(std::vector<double>& x) const { x = v; }
  std::size_t size() const { return v.size(); }
  double norm(std::string) const { double s = 0; for (double a : v) s += a * a; return std::sqrt(s); }
  std::vector<double> v;
};

class MultiMesh
{
public:
  std::size_t num_parts() const { return 0; }
  std::shared_ptr<const Mesh> part(std::size_t) const { return nullptr; }
};

class FiniteElement
{
public:
  explicit FiniteElement(std::shared_ptr<const ufc::finite_element> e) : ufc_element(e) {}
  std::shared_ptr<const ufc::finite_element> ufc_element;
};

class DofMap
{
public:
  DofMap(std::shared_ptr<const ufc::dofmap> d, const Mesh&) : ufc_dofmap(d) {}
  DofMap(std::shared_ptr<const ufc::dofmap> d, const Mesh&, std::shared_ptr<const SubDomain>) : ufc_dofmap(d) {}
  std::shared_ptr<const ufc::dofmap> ufc_dofmap;
};

// P1 ("Lagrange" degree 1, one dof per vertex, dof == vertex on one process), vector P1 and Real spaces are all
// the reference instantiates; only scalar P1 spaces carry Functions that are evaluated here.
class FunctionSpace
{
public:
  FunctionSpace(std::shared_ptr<const Mesh> mesh, std::shared_ptr<const FiniteElement> e, std::shared_ptr<const DofMap> d)
    : _mesh(mesh), _element(e), _dofmap(d) {}
  virtual ~FunctionSpace() {}
  std::shared_ptr<const Mesh> mesh() const { return _mesh; }
  std::shared_ptr<const FiniteElement> element() const { return _element; }
  std::size_t dim() const
  {
    const std::size_t per_cell = _element->ufc_element->space_dimension();
    const std::size_t nv = _mesh->tdim() + 1;
    return per_cell == 1 ? 1 : _mesh->num_vertices() * (per_cell / nv);
  }
private:
  std::shared_ptr<const Mesh> _mesh;
  std::shared_ptr<const FiniteElement> _element;
  std::shared_ptr<const DofMap> _dofmap;
};

inline std::vector<int> vertex_to_dof_map(const FunctionSpace& V)
{
  std::vector<int> m(V.mesh()->num_vertices());
  for (std::size_t i = 0; i < m.size(); ++i) m[i] = static_cast<int>(i);
  return m;
}
inline std::vector<std::size_t> dof_to_vertex_map(const FunctionSpace& V)
{
  std::vector<std::size_t> m(V.mesh()->num_vertices());
  for (std::size_t i = 0; i < m.size(); ++i) m[i] = i;
  return m;
}

class Function : public GenericFunction
{
public:
  explicit Function(std::shared_ptr<const FunctionSpace> V) : _V(V), _vec(std::make_shared<GenericVector>(V->dim())) {}
  std::shared_ptr<GenericVector> vector() { return _vec; }
  std::shared_ptr<const GenericVector> vector() const { return _vec; }
  std::shared_ptr<const FunctionSpace> function_space() const { return _V; }
  void set_allow_extrapolation(bool) {}
  // P1 interpolant at a point.  Interval meshes use x[0] only (a Function of the 1-D channel mesh is evaluated
  // at trap-wall points by the reference's DirichletBC); the rectangle mesh is uniform, so the cell is found by
  // index arithmetic.
  void eval(Array<double>& values, const Array<double>& x) const override
  {
    const Mesh& m = *_V->mesh();
    const std::vector<double>& X = m.coordinates();
    const std::vector<double>& u = _vec->v;
    if (m.gdim() == 1)
    {
      const std::size_t n = m.num_vertices();
      const double a = X[0], h = (X[n - 1] - X[0]) / static_cast<double>(n - 1);
      double t = (x[0] - a) / h;
      long i = static_cast<long>(std::floor(t));
      if (i < 0) i = 0;
      if (i > static_cast<long>(n) - 2) i = static_cast<long>(n) - 2;
      const double s = (x[0] - X[i]) / (X[i + 1] - X[i]);
      values[0] = u[i] * (1.0 - s) + u[i + 1] * s;
      return;
    }
    // 2-D: vertices row-major; nx+1 per row is recovered from the first cell's third vertex
    const std::size_t stride = m.cell(0)[2] - 1;   // cell 0 = (0, 1, nx+2): third vertex = TR = nx+2
    const std::size_t nxp = stride, nyp = m.num_vertices() / nxp;
    const double x0 = X[0], y0 = X[1], hx = X[2] - X[0], hy = X[2 * nxp + 1] - X[1];
    long ix = static_cast<long>(std::floor((x[0] - x0) / hx)), iy = static_cast<long>(std::floor((x[1] - y0) / hy));
    ix = std::max(0L, std::min(ix, static_cast<long>(nxp) - 2));
    iy = std::max(0L, std::min(iy, static_cast<long>(nyp) - 2));
    const double s = (x[0] - (x0 + ix * hx)) / hx, t = (x[1] - (y0 + iy * hy)) / hy;
    const std::size_t v0 = iy * nxp + ix, v1 = v0 + 1, v2 = v0 + nxp, v3 = v2 + 1;
    // lower triangle (v0, v1, v3) where s >= t, upper (v0, v2, v3) otherwise
    values[0] = s >= t ? u[v0] * (1.0 - s) + u[v1] * (s - t) + u[v3] * t : u[v0] * (1.0 - t) + u[v2] * (t - s) + u[v3] * s;
  }
  void interpolate(const GenericFunction& g)
  {
    const Mesh& m = *_V->mesh();
    const std::size_t gd = m.gdim();
    for (std::size_t v = 0; v < m.num_vertices(); ++v) _vec->v[v] = g.at(&m.coordinates()[v * gd], gd);
  }
private:
  std::shared_ptr<const FunctionSpace> _V;
  std::shared_ptr<GenericVector> _vec;
};

class MultiMeshFunctionSpace
{
public:
  explicit MultiMeshFunctionSpace(std::shared_ptr<const MultiMesh>) {}
  virtual ~MultiMeshFunctionSpace() {}
  void add(std::shared_ptr<const FunctionSpace>) {}
  void build() {}
  std::size_t num_parts() const { return 0; }
  std::shared_ptr<const FunctionSpace> part(std::size_t) const { return nullptr; }
  std::shared_ptr<const MultiMesh> multimesh() const { return nullptr; }
};

// ---- forms ----------------------------------------------------------------------------------------------------
class Form
{
public:
  Form(std::size_t rank, std::size_t ncoef) : _function_spaces(rank), _coefficients(ncoef) {}
  virtual ~Form() {}
  virtual std::size_t coefficient_number(const std::string&) const { return 0; }
  virtual std::string coefficient_name(std::size_t) const { return ""; }
  void set_mesh(std::shared_ptr<const Mesh> m) { _mesh = m; }
  void set_coefficient(std::size_t i, std::shared_ptr<const GenericFunction> c) { _coefficients.at(i) = c; }
  std::size_t rank() const { return _function_spaces.size(); }
  std::shared_ptr<const Mesh> mesh() const { return _function_spaces.empty() ? _mesh : _function_spaces[0]->mesh(); }
  std::shared_ptr<const FunctionSpace> function_space(std::size_t i) const { return _function_spaces[i]; }
  std::shared_ptr<const ufc::form> ufc_form() const { return _ufc_form; }
  const std::vector<std::shared_ptr<const GenericFunction>>& coefficients() const { return _coefficients; }
  // domain markers (DOLFIN: public shared_ptr members)
  std::shared_ptr<const MeshFunction<std::size_t>> dx, ds, dS, dP;
protected:
  std::vector<std::shared_ptr<const FunctionSpace>> _function_spaces;
  std::shared_ptr<const ufc::form> _ufc_form;
  std::shared_ptr<const Mesh> _mesh;
  std::vector<std::shared_ptr<const GenericFunction>> _coefficients;
};

class CoefficientAssigner
{
public:
  CoefficientAssigner(Form& f, std::size_t i) : _f(f), _i(i) {}
  const CoefficientAssigner& operator=(std::shared_ptr<const GenericFunction> c) { _f.set_coefficient(_i, c); return *this; }
private:
  Form& _f;
  std::size_t _i;
};

class MultiMeshForm
{
public:
  MultiMeshForm(std::shared_ptr<const MultiMeshFunctionSpace>, std::shared_ptr<const MultiMeshFunctionSpace>) {}
  explicit MultiMeshForm(std::shared_ptr<const MultiMeshFunctionSpace>) {}
  explicit MultiMeshForm(std::shared_ptr<const MultiMesh>) {}
  virtual ~MultiMeshForm() {}
  void add(std::shared_ptr<const Form>) {}
  void build() {}
};
class MultiMeshCoefficientAssigner
{
public:
  MultiMeshCoefficientAssigner(MultiMeshForm&, std::size_t) {}
  const MultiMeshCoefficientAssigner& operator=(std::shared_ptr<const GenericFunction>) { return *this; }
};

// ---- assembly ---------------------------------------------------------------------------------------------------
// Sparse rows as sorted maps: the default trap has 8 241 unknowns with <= 7 entries per row.
struct SparseSystem
{
  std::vector<std::map<std::size_t, double>> A;
  std::vector<double> b;
};

namespace shim_detail
{
// coefficient values restricted to a cell: one value for a Real coefficient, vertex values for P1
inline void restrict_coefficients(const Form& form, const Mesh& mesh, const std::size_t* cv, std::vector<std::vector<double>>& w,
                                  std::vector<const double*>& wp)
{
  const std::size_t nc = form.coefficients().size(), nv = mesh.tdim() + 1, gd = mesh.gdim(), rank = form.rank();
  w.resize(nc); wp.resize(nc);
  for (std::size_t i = 0; i < nc; ++i)
  {
    std::unique_ptr<ufc::finite_element> e(form.ufc_form()->create_finite_element(rank + i));
    const std::size_t nd = e->space_dimension();
    w[i].assign(nd, 0.0);
    const GenericFunction* g = form.coefficients()[i].get();
    if (!g) dolfin_error("Assembler", "restrict coefficient", "coefficient " + form.coefficient_name(i) + " has not been set");
    if (nd == 1) w[i][0] = g->at(&mesh.coordinates()[cv[0] * gd], gd);
    else
    {
      const Function* fn = dynamic_cast<const Function*>(g);
      for (std::size_t k = 0; k < nv; ++k)
        w[i][k] = (fn && fn->function_space()->mesh().get() == &mesh) ? fn->vector()->v[cv[k]]
                                                                      : g->at(&mesh.coordinates()[cv[k] * gd], gd);
    }
    wp[i] = w[i].data();
  }
}
inline void cell_coordinates(const Mesh& mesh, const std::size_t* cv, std::vector<double>& xy)
{
  const std::size_t nv = mesh.tdim() + 1, gd = mesh.gdim();
  xy.resize(nv * gd);
  for (std::size_t k = 0; k < nv; ++k)
    for (std::size_t d = 0; d < gd; ++d) xy[k * gd + d] = mesh.coordinates()[cv[k] * gd + d];
}
}

// Assembles a rank-2, rank-1 or rank-0 form: cells with the default cell integral, boundary facets with the
// exterior-facet integral of their marker (forms' ds) or the default one.
inline void assemble_form(const Form& form, SparseSystem* sys, bool matrix, double* scalar)
{
  const Mesh& mesh = *form.mesh();
  const std::size_t nv = mesh.tdim() + 1, rank = form.rank();
  const ufc::form& uf = *form.ufc_form();
  std::vector<std::vector<double>> w;
  std::vector<const double*> wp;
  std::vector<double> xy, Ae(rank == 2 ? nv * nv : (rank == 1 ? nv : 1));
  auto scatter = [&](const std::size_t* cv) {
    if (rank == 2) { for (std::size_t i = 0; i < nv; ++i) for (std::size_t j = 0; j < nv; ++j) sys->A[cv[i]][cv[j]] += Ae[i * nv + j]; }
    else if (rank == 1) { for (std::size_t i = 0; i < nv; ++i) sys->b[cv[i]] += Ae[i]; }
    else *scalar += Ae[0];
  };
  (void)matrix;
  if (uf.has_cell_integrals())
  {
    std::unique_ptr<ufc::cell_integral> ci(uf.create_default_cell_integral());
    if (ci)
      for (std::size_t c = 0; c < mesh.num_cells(); ++c)
      {
        const std::size_t* cv = mesh.cell(c);
        shim_detail::restrict_coefficients(form, mesh, cv, w, wp);
        shim_detail::cell_coordinates(mesh, cv, xy);
        std::fill(Ae.begin(), Ae.end(), 0.0);
        ci->tabulate_tensor(Ae.data(), wp.data(), xy.data(), 0);
        scatter(cv);
      }
  }
  if (uf.has_exterior_facet_integrals())
  {
    std::unique_ptr<ufc::exterior_facet_integral> dflt(uf.create_default_exterior_facet_integral());
    std::map<std::size_t, std::unique_ptr<ufc::exterior_facet_integral>> by_id;
    const auto& bf = mesh.boundary_facets();
    for (std::size_t k = 0; k < bf.size(); ++k)
    {
      const ufc::exterior_facet_integral* fi = dflt.get();
      if (form.ds)
      {
        const std::size_t id = form.ds->boundary_value(k);
        if (id < uf.max_exterior_facet_subdomain_id())
        {
          if (!by_id.count(id)) by_id[id].reset(uf.create_exterior_facet_integral(id));
          if (by_id[id]) fi = by_id[id].get();
        }
      }
      if (!fi) continue;
      const std::size_t* cv = mesh.cell(bf[k].cell);
      shim_detail::restrict_coefficients(form, mesh, cv, w, wp);
      shim_detail::cell_coordinates(mesh, cv, xy);
      std::fill(Ae.begin(), Ae.end(), 0.0);
      fi->tabulate_tensor(Ae.data(), wp.data(), xy.data(), bf[k].local, 0);
      scatter(cv);
    }
  }
}

class DirichletBC
{
public:
  DirichletBC(std::shared_ptr<const FunctionSpace> V, std::shared_ptr<const GenericFunction> g, std::shared_ptr<const SubDomain> sub)
    : _V(V), _g(g), _sub(sub)
  {
    // the reference reaches this with a null sub-domain for DIRICHLET_0 + a trap type its four `if`s do not name
    // (src/fHSL.cpp:559-566, e.g. H_TRAP): undefined behaviour upstream, a clean error here
    if (!sub) dolfin_error("DirichletBC", "create boundary condition", "null sub-domain");
  }
  // "topological" search: dofs of the boundary facets that lie inside the sub-domain, values g(x_dof)
  void get_boundary_values(std::map<std::size_t, double>& bv) const
  {
    const Mesh& mesh = *_V->mesh();
    const std::size_t gd = mesh.gdim();
    for (const BoundaryFacet& f : mesh.boundary_facets())
      if (_sub->facet_inside(mesh, f))
        for (std::size_t v : f.v) bv[v] = _g->at(&mesh.coordinates()[v * gd], gd);
  }
private:
  std::shared_ptr<const FunctionSpace> _V;
  std::shared_ptr<const GenericFunction> _g;
  std::shared_ptr<const SubDomain> _sub;
};

class LinearVariationalProblem
{
public:
  LinearVariationalProblem(std::shared_ptr<const Form> a, std::shared_ptr<const Form> L, std::shared_ptr<Function> u,
                           std::vector<std::shared_ptr<const DirichletBC>> bcs)
    : a(a), L(L), u(u), bcs(bcs) {}
  std::shared_ptr<const Form> a, L;
  std::shared_ptr<Function> u;
  std::vector<std::shared_ptr<const DirichletBC>> bcs;
};

// solve(): assemble a and L, apply the DirichletBCs in list order (identity rows, columns kept -- the
// non-symmetric path of DOLFIN's SystemAssembler-free LinearVariationalSolver with symmetric = false), banded LU.
class LinearVariationalSolver
{
public:
  explicit LinearVariationalSolver(std::shared_ptr<LinearVariationalProblem> p) : _p(p) {}
  void solve()
  {
    const std::size_t n = _p->u->vector()->size();
    SparseSystem sys;
    sys.A.assign(n, {});
    sys.b.assign(n, 0.0);
    assemble_form(*_p->a, &sys, true, nullptr);
    SparseSystem rhs;
    rhs.b.assign(n, 0.0);
    assemble_form(*_p->L, &rhs, false, nullptr);
    sys.b = rhs.b;
    for (const auto& bc : _p->bcs)
    {
      std::map<std::size_t, double> bv;
      bc->get_boundary_values(bv);
      for (const auto& kv : bv)
      {
        sys.A[kv.first].clear();
        sys.A[kv.first][kv.first] = 1.0;
        sys.b[kv.first] = kv.second;
      }
    }
    // banded LU without pivoting
    std::size_t bw = 0;
    for (std::size_t i = 0; i < n; ++i)
      for (const auto& e : sys.A[i]) bw = std::max(bw, e.first > i ? e.first - i : i - e.first);
    const std::size_t W = 2 * bw + 1;
    std::vector<double> B(n * W, 0.0);   // B[i*W + (j - i + bw)]
    for (std::size_t i = 0; i < n; ++i)
      for (const auto& e : sys.A[i]) B[i * W + (e.first + bw - i)] = e.second;
    std::vector<double> x = sys.b;
    for (std::size_t k = 0; k < n; ++k)
    {
      const double piv = B[k * W + bw];
      const std::size_t iend = std::min(n, k + bw + 1);
      for (std::size_t i = k + 1; i < iend; ++i)
      {
        const double lik = B[i * W + (k + bw - i)];
        if (lik == 0.0) continue;
        const double m = lik / piv;
        const std::size_t jend = std::min(n, k + bw + 1);
        for (std::size_t j = k + 1; j < jend; ++j) B[i * W + (j + bw - i)] -= m * B[k * W + (j + bw - k)];
        x[i] -= m * x[k];
      }
    }
    for (std::size_t kk = n; kk-- > 0;)
    {
      double s = x[kk];
      const std::size_t jend = std::min(n, kk + bw + 1);
      for (std::size_t j = kk + 1; j < jend; ++j) s -= B[kk * W + (j + bw - kk)] * x[j];
      x[kk] = s / B[kk * W + bw];
    }
    _p->u->vector()->v = x;
  }
private:
  std::shared_ptr<LinearVariationalProblem> _p;
};

class Scalar
{
public:
  explicit Scalar(MPI_Comm = 0) : value(0.0) {}
  double get_scalar_value() const { return value; }
  double value;
};

class Assembler
{
public:
  void assemble(Scalar& s, const Form& form)
  {
    s.value = 0.0;
    assemble_form(form, nullptr, false, &s.value);
  }
};

class File
{
public:
  File(MPI_Comm, const std::string&, const std::string& = "") {}
  File(const std::string&, const std::string& = "") {}
  template <class T> File& operator<<(const T&) { return *this; }
};
}
#endif
