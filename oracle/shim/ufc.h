// Minimal stand-in for <ufc.h> (UFC 2018.1.0 interface) so that the reference's
// FFC-generated headers (/root/reference/fenics/*.h) compile IN PLACE without
// FEniCS installed.  TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Written for
// this repository: it declares only the pure-virtual interface the generated
// classes override (signatures taken from their `final override` declarations)
// plus the two reference-cell tables the facet kernels index.  No FEniCS code.
#ifndef EQ_B200_SHIM_UFC_H
#define EQ_B200_SHIM_UFC_H
#include <cstddef>
#include <vector>

namespace ufc
{
enum class shape { interval, triangle, quadrilateral, tetrahedron, hexahedron, vertex };

class cell
{
public:
  virtual ~cell() {}
  shape cell_shape;
  std::size_t topological_dimension;
  std::size_t geometric_dimension;
  std::vector<std::vector<std::size_t>> entity_indices;
  std::size_t index;
  int local_facet;
  int orientation;
  int mesh_identifier;
};

class function
{
public:
  virtual ~function() {}
  virtual void evaluate(double * values, const double * coordinates, const cell& c) const = 0;
};

class coordinate_mapping;
class dofmap;

class finite_element
{
public:
  virtual ~finite_element() {}
  virtual const char * signature() const = 0;
  virtual shape cell_shape() const = 0;
  virtual std::size_t topological_dimension() const = 0;
  virtual std::size_t geometric_dimension() const = 0;
  virtual std::size_t space_dimension() const = 0;
  virtual std::size_t value_rank() const = 0;
  virtual std::size_t value_dimension(std::size_t i) const = 0;
  virtual std::size_t value_size() const = 0;
  virtual std::size_t reference_value_rank() const = 0;
  virtual std::size_t reference_value_dimension(std::size_t i) const = 0;
  virtual std::size_t reference_value_size() const = 0;
  virtual std::size_t degree() const = 0;
  virtual const char * family() const = 0;
  virtual void evaluate_reference_basis(double * reference_values, std::size_t num_points, const double * X) const = 0;
  virtual void evaluate_reference_basis_derivatives(double * reference_values, std::size_t order, std::size_t num_points, const double * X) const = 0;
  virtual void transform_reference_basis_derivatives(double * values, std::size_t order, std::size_t num_points, const double * reference_values, const double * X, const double * J, const double * detJ, const double * K, int cell_orientation) const = 0;
  virtual void evaluate_basis(std::size_t i, double * values, const double * x, const double * coordinate_dofs, int cell_orientation, const coordinate_mapping * cm ) const = 0;
  virtual void evaluate_basis_all(double * values, const double * x, const double * coordinate_dofs, int cell_orientation, const coordinate_mapping * cm ) const = 0;
  virtual void evaluate_basis_derivatives(std::size_t i, std::size_t n, double * values, const double * x, const double * coordinate_dofs, int cell_orientation, const coordinate_mapping * cm ) const = 0;
  virtual void evaluate_basis_derivatives_all(std::size_t n, double * values, const double * x, const double * coordinate_dofs, int cell_orientation, const coordinate_mapping * cm ) const = 0;
  virtual double evaluate_dof(std::size_t i, const function& f, const double * coordinate_dofs, int cell_orientation, const cell& c, const coordinate_mapping * cm ) const = 0;
  virtual void evaluate_dofs(double * values, const function& f, const double * coordinate_dofs, int cell_orientation, const cell& c, const coordinate_mapping * cm ) const = 0;
  virtual void interpolate_vertex_values(double * vertex_values, const double * dof_values, const double * coordinate_dofs, int cell_orientation, const coordinate_mapping * cm ) const = 0;
  virtual void tabulate_dof_coordinates(double * dof_coordinates, const double * coordinate_dofs, const coordinate_mapping * cm ) const = 0;
  virtual void tabulate_reference_dof_coordinates(double * reference_dof_coordinates) const = 0;
  virtual std::size_t num_sub_elements() const = 0;
  virtual finite_element * create_sub_element(std::size_t i) const = 0;
  virtual finite_element * create() const = 0;
};

class dofmap
{
public:
  virtual ~dofmap() {}
  virtual const char * signature() const = 0;
  virtual bool needs_mesh_entities(std::size_t d) const = 0;
  virtual std::size_t topological_dimension() const = 0;
  virtual std::size_t global_dimension(const std::vector<std::size_t>& num_global_entities) const = 0;
  virtual std::size_t num_global_support_dofs() const = 0;
  virtual std::size_t num_element_support_dofs() const = 0;
  virtual std::size_t num_element_dofs() const = 0;
  virtual std::size_t num_facet_dofs() const = 0;
  virtual std::size_t num_entity_dofs(std::size_t d) const = 0;
  virtual std::size_t num_entity_closure_dofs(std::size_t d) const = 0;
  virtual void tabulate_dofs(std::size_t * dofs, const std::vector<std::size_t>& num_global_entities, const std::vector<std::vector<std::size_t>>& entity_indices) const = 0;
  virtual void tabulate_facet_dofs(std::size_t * dofs, std::size_t facet) const = 0;
  virtual void tabulate_entity_dofs(std::size_t * dofs, std::size_t d, std::size_t i) const = 0;
  virtual void tabulate_entity_closure_dofs(std::size_t * dofs, std::size_t d, std::size_t i) const = 0;
  virtual std::size_t num_sub_dofmaps() const = 0;
  virtual dofmap * create_sub_dofmap(std::size_t i) const = 0;
  virtual dofmap * create() const = 0;
};

class coordinate_mapping
{
public:
  virtual ~coordinate_mapping() {}
  virtual const char * signature() const = 0;
  virtual coordinate_mapping * create() const = 0;
  virtual std::size_t geometric_dimension() const = 0;
  virtual std::size_t topological_dimension() const = 0;
  virtual shape cell_shape() const = 0;
  virtual finite_element * create_coordinate_finite_element() const = 0;
  virtual dofmap * create_coordinate_dofmap() const = 0;
  virtual void compute_physical_coordinates( double * x, std::size_t num_points, const double * X, const double * coordinate_dofs) const = 0;
  virtual void compute_reference_coordinates( double * X, std::size_t num_points, const double * x, const double * coordinate_dofs, int cell_orientation) const = 0;
  virtual void compute_reference_geometry( double * X, double * J, double * detJ, double * K, std::size_t num_points, const double * x, const double * coordinate_dofs, int cell_orientation) const = 0;
  virtual void compute_jacobians( double * J, std::size_t num_points, const double * X, const double * coordinate_dofs) const = 0;
  virtual void compute_jacobian_determinants( double * detJ, std::size_t num_points, const double * J, int cell_orientation) const = 0;
  virtual void compute_jacobian_inverses( double * K, std::size_t num_points, const double * J, const double * detJ) const = 0;
  virtual void compute_geometry( double * x, double * J, double * detJ, double * K, std::size_t num_points, const double * X, const double * coordinate_dofs, int cell_orientation) const = 0;
  virtual void compute_midpoint_geometry( double * x, double * J, const double * coordinate_dofs) const = 0;
};

class cell_integral
{
public:
  virtual ~cell_integral() {}
  virtual const std::vector<bool> & enabled_coefficients() const = 0;
  virtual void tabulate_tensor(double * A, const double * const * w, const double * coordinate_dofs, int cell_orientation) const = 0;
};

class exterior_facet_integral
{
public:
  virtual ~exterior_facet_integral() {}
  virtual const std::vector<bool> & enabled_coefficients() const = 0;
  virtual void tabulate_tensor(double * A, const double * const * w, const double * coordinate_dofs, std::size_t facet, int cell_orientation) const = 0;
};

class interior_facet_integral
{
public:
  virtual ~interior_facet_integral() {}
  virtual const std::vector<bool> & enabled_coefficients() const = 0;
};

class vertex_integral
{
public:
  virtual ~vertex_integral() {}
  virtual const std::vector<bool> & enabled_coefficients() const = 0;
};

class custom_integral
{
public:
  virtual ~custom_integral() {}
  virtual const std::vector<bool> & enabled_coefficients() const = 0;
};

class cutcell_integral
{
public:
  virtual ~cutcell_integral() {}
  virtual const std::vector<bool> & enabled_coefficients() const = 0;
};

class interface_integral
{
public:
  virtual ~interface_integral() {}
  virtual const std::vector<bool> & enabled_coefficients() const = 0;
};

class overlap_integral
{
public:
  virtual ~overlap_integral() {}
  virtual const std::vector<bool> & enabled_coefficients() const = 0;
};

class form
{
public:
  virtual ~form() {}
  virtual const char * signature() const = 0;
  virtual std::size_t rank() const = 0;
  virtual std::size_t num_coefficients() const = 0;
  virtual std::size_t original_coefficient_position(std::size_t i) const = 0;
  virtual finite_element * create_coordinate_finite_element() const = 0;
  virtual dofmap * create_coordinate_dofmap() const = 0;
  virtual coordinate_mapping * create_coordinate_mapping() const = 0;
  virtual finite_element * create_finite_element(std::size_t i) const = 0;
  virtual dofmap * create_dofmap(std::size_t i) const = 0;
  virtual std::size_t max_cell_subdomain_id() const = 0;
  virtual std::size_t max_exterior_facet_subdomain_id() const = 0;
  virtual std::size_t max_interior_facet_subdomain_id() const = 0;
  virtual std::size_t max_vertex_subdomain_id() const = 0;
  virtual std::size_t max_custom_subdomain_id() const = 0;
  virtual std::size_t max_cutcell_subdomain_id() const = 0;
  virtual std::size_t max_interface_subdomain_id() const = 0;
  virtual std::size_t max_overlap_subdomain_id() const = 0;
  virtual bool has_cell_integrals() const = 0;
  virtual bool has_exterior_facet_integrals() const = 0;
  virtual bool has_interior_facet_integrals() const = 0;
  virtual bool has_vertex_integrals() const = 0;
  virtual bool has_custom_integrals() const = 0;
  virtual bool has_cutcell_integrals() const = 0;
  virtual bool has_interface_integrals() const = 0;
  virtual bool has_overlap_integrals() const = 0;
  virtual cell_integral * create_cell_integral(std::size_t subdomain_id) const = 0;
  virtual exterior_facet_integral * create_exterior_facet_integral(std::size_t subdomain_id) const = 0;
  virtual interior_facet_integral * create_interior_facet_integral(std::size_t subdomain_id) const = 0;
  virtual vertex_integral * create_vertex_integral(std::size_t subdomain_id) const = 0;
  virtual custom_integral * create_custom_integral(std::size_t subdomain_id) const = 0;
  virtual cutcell_integral * create_cutcell_integral(std::size_t subdomain_id) const = 0;
  virtual interface_integral * create_interface_integral(std::size_t subdomain_id) const = 0;
  virtual overlap_integral * create_overlap_integral(std::size_t subdomain_id) const = 0;
  virtual cell_integral * create_default_cell_integral() const = 0;
  virtual exterior_facet_integral * create_default_exterior_facet_integral() const = 0;
  virtual interior_facet_integral * create_default_interior_facet_integral() const = 0;
  virtual vertex_integral * create_default_vertex_integral() const = 0;
  virtual custom_integral * create_default_custom_integral() const = 0;
  virtual cutcell_integral * create_default_cutcell_integral() const = 0;
  virtual interface_integral * create_default_interface_integral() const = 0;
  virtual overlap_integral * create_default_overlap_integral() const = 0;
};

} // namespace ufc

// Reference-cell geometry tables of ufc_geometry.h that the generated facet
// kernels index: facet f is the edge opposite reference vertex f.
static const double interval_reference_facet_normals[2][1] = { { -1.0 }, { +1.0 } };
static const double triangle_reference_facet_jacobian[3][2][1] = {
  { { -1.0 }, { 1.0 } }, { { 0.0 }, { 1.0 } }, { { 1.0 }, { 0.0 } } };
static const double triangle_reference_facet_normals[3][2] = {
  { 0.7071067811865476, 0.7071067811865476 }, { -1.0, 0.0 }, { 0.0, -1.0 } };

// Affine-geometry helpers of ufc_geometry.h used by the generated
// finite_element::evaluate_basis* bodies (not on the oracle's path).
inline void compute_jacobian_interval_1d(double* J, const double* cd) { J[0] = cd[1] - cd[0]; }
inline void compute_jacobian_inverse_interval_1d(double* K, double& det, const double* J)
{ det = J[0]; K[0] = 1.0 / det; }
inline void compute_jacobian_triangle_2d(double* J, const double* cd)
{ J[0] = cd[2] - cd[0]; J[1] = cd[4] - cd[0]; J[2] = cd[3] - cd[1]; J[3] = cd[5] - cd[1]; }
inline void compute_jacobian_inverse_triangle_2d(double* K, double& det, const double* J)
{ det = J[0] * J[3] - J[1] * J[2]; K[0] = J[3] / det; K[1] = -J[1] / det; K[2] = -J[2] / det; K[3] = J[0] / det; }

#endif
