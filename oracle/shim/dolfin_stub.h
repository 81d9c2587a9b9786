// Empty-bodied stand-ins for the handful of DOLFIN classes that the trailing
// "-l dolfin" wrapper section of the FFC-generated headers derives from, so the
// headers compile unmodified where they lie.  TEST INFRASTRUCTURE ONLY; nothing
// here does any work -- the oracle/_ref build only calls the generated
// ufc::*_integral::tabulate_tensor bodies.  Written for this repository.
#ifndef EQ_B200_SHIM_DOLFIN_STUB_H
#define EQ_B200_SHIM_DOLFIN_STUB_H
#include <cstddef>
#include <memory>
#include <string>
#include <vector>
#include <ufc.h>

namespace dolfin
{
class Mesh {};
class SubDomain {};
class GenericFunction {};
class GenericVector {};

class MultiMesh
{
public:
  std::size_t num_parts() const { return 0; }
  std::shared_ptr<const Mesh> part(std::size_t) const { return nullptr; }
};

class FiniteElement
{
public:
  explicit FiniteElement(std::shared_ptr<const ufc::finite_element>) {}
};

class DofMap
{
public:
  DofMap(std::shared_ptr<const ufc::dofmap>, const Mesh&) {}
  DofMap(std::shared_ptr<const ufc::dofmap>, const Mesh&, std::shared_ptr<const SubDomain>) {}
};

class FunctionSpace
{
public:
  FunctionSpace(std::shared_ptr<const Mesh>, std::shared_ptr<const FiniteElement>,
                std::shared_ptr<const DofMap>) {}
  virtual ~FunctionSpace() {}
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

class Form
{
public:
  Form(std::size_t rank, std::size_t) : _function_spaces(rank) {}
  virtual ~Form() {}
  virtual std::size_t coefficient_number(const std::string&) const { return 0; }
  virtual std::string coefficient_name(std::size_t) const { return ""; }
  void set_mesh(std::shared_ptr<const Mesh> m) { _mesh = m; }
protected:
  std::vector<std::shared_ptr<const FunctionSpace>> _function_spaces;
  std::shared_ptr<const ufc::form> _ufc_form;
  std::shared_ptr<const Mesh> _mesh;
};

class CoefficientAssigner
{
public:
  CoefficientAssigner(Form&, std::size_t) {}
  const CoefficientAssigner& operator=(std::shared_ptr<const GenericFunction>) { return *this; }
};

class MultiMeshForm
{
public:
  MultiMeshForm(std::shared_ptr<const MultiMeshFunctionSpace>,
                std::shared_ptr<const MultiMeshFunctionSpace>) {}
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

inline void dolfin_error(std::string, std::string, std::string, ...) {}
}
#endif
