// gsrb_fused.cu -- fused red+black GSRB sweep (placeholder: sequences the per-colour kernel until the
// plane-streaming kernel lands; same results by construction).
#include "mgic_internal.h"

namespace mgk {
int gsrb_fused(mgic_op *o, mgic_field *e, const mgic_field *r, int iterations) {
  const Geom g = o->geom();
  const BCk bc = o->bck(true);
  for (int it = 0; it < iterations; it++)
    for (int pass = 0; pass <= 1; pass++)
      MGIC_TRY(gsrb_color(o->ctx, g, bc, e->p, r->p, o->a->p, o->b ? o->b->p : nullptr, o->lambda->p, o->alpha, o->beta, o->dx, pass));
  return MGIC_OK;
}
}  // namespace mgk
