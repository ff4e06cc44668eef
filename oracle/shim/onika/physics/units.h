#pragma once
namespace onika { namespace physics { struct Quantity { double v = 0; double convert() const { return v; } }; } }
