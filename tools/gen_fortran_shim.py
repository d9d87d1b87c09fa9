#!/usr/bin/env python
"""Generates the Fortran side of the boundary from the single sources of truth:

  include/mpasb.h          -> fortran/mpasb_binding.F90   (ISO_C_BINDING interface of every exported symbol, both RKIND widths)
  include/mpasb_fields.def -> fortran/mpas_atm_dynamics_b200.F  (replacement bodies of mpas_atm_dynamics_init / _finalize /
                              atm_timestep with every upload and download spelled out by pool and key)

The pool of each key is the var_struct it sits in in src/core_atmosphere/Registry.xml (state / diag / tend / tend_physics /
mesh; `name_in_code` where it differs from the registry name); module scratch of mpas_atm_time_integration.F:90-140 never
crosses the boundary.  tests/test_abi.py regenerates both files and compares them with the committed ones."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# key -> (pool, name in the pool).  From Registry.xml; module scratch arrays are absent on purpose.
TEND = {"tend_u": "u", "tend_w": "w", "tend_theta": "theta_m", "tend_rho": "rho_zz", "rt_diabatic_tend": "rt_diabatic_tend",
        "scalars_tend": "scalars_tend"}
TEND_PHYSICS = ("rthdynten", "tend_ru_physics", "tend_rtheta_physics", "tend_rho_physics")
STATE = ("u", "w", "rho_zz", "theta_m", "scalars")
LBC = ("lbc_u", "lbc_ru", "lbc_rho_zz", "lbc_rtheta_m", "lbc_scalars")      # var_struct "lbc", two time levels
SCRATCH = ("tend_u_euler", "tend_w_euler", "tend_theta_euler", "qtot", "delsq_theta", "delsq_w", "delsq_divergence", "dpdz", "delsq_u",
           "delsq_vorticity", "ke_vertex", "ke_edge", "scalar_old", "scalar_new", "s_max", "s_min", "rho_zz_int", "wdtn", "flux_arr",
           "flux_upwind_tmp", "flux_tmp", "scale_arr", "horiz_flux_arr")
# diag fields that carry information INTO a step (restart stream + what the init-time routines leave behind)
DIAG_IN = ("ru", "rw", "rtheta_p", "rho_p", "rho_base", "rtheta_base", "theta_base", "exner", "exner_base", "pressure_p", "pressure_base",
           "pv_edge", "rho_edge", "v", "ke", "divergence", "vorticity", "pv_vertex", "pv_cell", "gradPVn", "gradPVt", "h_divergence",
           "kdiff", "cqw", "cqu", "ruAvg", "wwAvg", "ruAvg_split", "wwAvg_split", "rho_zz_old_split", "theta", "rho")
# what an output / restart alarm needs back (Registry.xml streams "output" and "restart")
DIAG_OUT = ("ru", "rw", "rtheta_p", "rho_p", "exner", "pressure_p", "pv_edge", "rho_edge", "v", "ke", "divergence", "vorticity",
            "pv_vertex", "uReconstructX", "uReconstructY", "uReconstructZ", "uReconstructZonal", "uReconstructMeridional",
            "theta", "rho", "pressure", "ruAvg", "wwAvg")


def fields():
    out = []
    txt = open(os.path.join(ROOT, "include", "mpasb_fields.def")).read()
    for m in re.finditer(r"^F\((\w+),\s*(\w+),\s*(\w+),\s*(\d),\s*(\w+),\s*(\w+)\)", txt, re.M):
        out.append(dict(name=m.group(1), loc=m.group(2), inner=m.group(3), levels=int(m.group(4)), type=m.group(5)))
    return out


def rank_of(f):
    if f["loc"] == "LEVS":
        return 1
    return {"ONE": 1, "NL": 2, "NL1": 2, "ME": 2, "ME2": 2, "VD": 2, "TWO": 2, "F15": 2, "NL1_ME": 3, "S_NL": 3, "NL_TWO": 3, "THREE_ME": 3}[f["inner"]]


def struct_members(header, name):
    body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if stmt:
            ctype, names = stmt.split(None, 1)
            out += [(ctype, n.strip()) for n in names.split(",")]
    return out


CT = {"int": "integer(c_int), value", "long": "integer(c_long), value", "mpasb_real": "real(mpasb_real), value", "double": "real(c_double), value"}


def f_arg(ctype, name):
    """One C parameter -> (Fortran dummy name, declaration)."""
    ctype = ctype.strip()
    if ctype == "mpasb_handle":
        return name, f"type(c_ptr), value :: {name}"
    if ctype == "mpasb_handle*":
        return name, f"type(c_ptr), intent(out) :: {name}"
    if ctype in ("const mpasb_dims*", "const mpasb_config*"):
        return name, f"type({ctype.split()[1][:-1]}), intent(in) :: {name}"
    if ctype == "const char*":
        return name, f"character(kind=c_char), intent(in) :: {name}(*)"
    if ctype == "char*":
        return name, f"character(kind=c_char), intent(out) :: {name}(*)"
    if ctype == "const char* const*":
        return name, f"type(c_ptr), intent(in) :: {name}(*)        ! C strings: c_loc of null-terminated character arrays"
    if ctype in ("const mpasb_real* const*", "mpasb_real* const*", "const void* const*", "void* const*"):
        return name, f"type(c_ptr), intent(in) :: {name}(*)        ! c_loc of the host arrays"
    if ctype in ("const void*", "void*"):
        return name, f"character(kind=c_char) :: {name}(*)"
    m = re.match(r"^(const )?(int|long|double|mpasb_real)\*$", ctype)
    if m:
        base = {"int": "integer(c_int)", "long": "integer(c_long)", "double": "real(c_double)", "mpasb_real": "real(mpasb_real)"}[m.group(2)]
        return name, f"{base}, intent({'in' if m.group(1) else 'inout'}) :: {name}(*)"
    return name, f"{CT[ctype]} :: {name}"


def prototypes(header):
    txt = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    out = []
    for m in re.finditer(r"^(int|long|const char\*)\s+(mpasb_\w+)\(([^;]*?)\);", txt, re.M | re.S):
        ret, name, params = m.group(1), m.group(2), " ".join(m.group(3).split())
        args = []
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                arr = re.match(r"^(.*?)(\w+)\[\d*\]$", p)
                if arr:
                    args.append((arr.group(1).strip() + "*", arr.group(2)))
                else:
                    mm = re.match(r"^(.*?)(\w+)$", p)
                    args.append((mm.group(1).strip(), mm.group(2)))
        out.append((ret, name, args))
    return out


def binding():
    header = open(os.path.join(ROOT, "include", "mpasb.h")).read()
    L = ["! mpasb_binding.F90 -- ISO_C_BINDING interface of libmpasb.so / libmpasb_sp.so (include/mpasb.h).",
         "! GENERATED by tools/gen_fortran_shim.py from the header; do not edit.  Compile with the reference's own flags",
         "! (-DSINGLE_PRECISION selects RKIND = single, reference Makefile:861-873) and add it to",
         "! src/core_atmosphere/dynamics/Makefile ahead of mpas_atm_time_integration.o.",
         "module mpasb_binding", "   use iso_c_binding", "   implicit none", "   public", "",
         "#ifdef SINGLE_PRECISION", "   integer, parameter :: mpasb_real = c_float      ! libmpasb_sp.so", "#else",
         "   integer, parameter :: mpasb_real = c_double     ! libmpasb.so", "#endif", ""]
    for sname in ("mpasb_dims", "mpasb_config"):
        L.append(f"   type, bind(C) :: {sname}")
        for ctype, n in struct_members(header, sname):
            L.append(f"      {'integer(c_int)' if ctype == 'int' else 'real(c_double)'} :: {n}")
        L.append(f"   end type {sname}")
        L.append("")
    L.append("   interface")
    for ret, name, args in prototypes(header):
        fret = {"int": "integer(c_int)", "long": "integer(c_long)", "const char*": "type(c_ptr)"}[ret]
        names = ", ".join(a[1] for a in args)
        L.append(f"      {fret} function {name}({names}) bind(C, name='{name}')")
        L.append("         import")
        for ctype, n in args:
            L.append("         " + f_arg(ctype, n)[1])
        L.append(f"      end function {name}")
    L += ["   end interface", "", "contains", "",
          "   ! C string (mpasb_last_error) -> Fortran string",
          "   function mpasb_c_to_f(cstr) result(s)", "      type(c_ptr), intent(in) :: cstr", "      character(len=:), allocatable :: s",
          "      character(kind=c_char), pointer :: p(:)", "      integer :: n",
          "      s = ''", "      if (.not. c_associated(cstr)) return", "      call c_f_pointer(cstr, p, [4096])", "      n = 0",
          "      do while (n < 4096)", "         if (p(n+1) == c_null_char) exit", "         n = n + 1", "      end do",
          "      allocate(character(len=n) :: s)", "      s = transfer(p(1:n), s)", "   end function mpasb_c_to_f", "",
          "end module mpasb_binding", ""]
    return "\n".join(L)


def pool_of(name):
    if name in STATE:
        return "state", name
    if name in TEND:
        return "tend", TEND[name]
    if name in TEND_PHYSICS:
        return "tend_physics", name
    return None, name


def dynamics():
    F = [f for f in fields() if f["name"] not in SCRATCH]
    mesh = [f for f in F if pool_of(f["name"])[0] is None and f["name"] in MESH_KEYS(F)]
    diag = [f for f in F if pool_of(f["name"])[0] is None and f["name"] not in MESH_KEYS(F)]
    byname = {f["name"]: f for f in F}
    L = []
    A = L.append
    A("! mpas_atm_dynamics_b200.F -- replacement bodies of the three entry points of the dycore step, calling libmpasb.so")
    A("! through mpasb_binding.  GENERATED by tools/gen_fortran_shim.py from include/mpasb_fields.def; do not edit.")
    A("!")
    A("! Drop-in: in src/core_atmosphere/dynamics/mpas_atm_time_integration.F rename the reference's mpas_atm_dynamics_init,")
    A("! mpas_atm_dynamics_finalize and atm_timestep (TI:205, 479, 739) or build with -DMPASB_DYCORE and `use` this module instead;")
    A("! names, argument lists and the call sites mpas_atm_core.F:296, 1022, 1054 stay as they are.  Pools, streams, namelists and")
    A("! the driver are untouched: host arrays are import/export only, looked up by the same pool keys every time (never cached:")
    A("! time levels rotate by pointer swap, shift_time_levs_array.inc:26-33).")
    A("module atm_time_integration_b200")
    A("")
    for u in ("iso_c_binding", "mpas_derived_types", "mpas_pool_routines", "mpas_kind_types", "mpas_constants", "mpas_dmpar", "mpas_log",
              "mpas_timer", "mpas_timekeeping", "mpas_atm_boundaries", "mpasb_binding"):
        A(f"   use {u}")
    A("")
    A("   implicit none")
    A("   private")
    A("   public :: mpas_atm_dynamics_init, mpas_atm_dynamics_finalize, atm_timestep, mpasb_download_for_output")
    A("")
    A("   type(c_ptr), save :: mpasb_h = c_null_ptr")
    A("   logical, save :: state_on_device = .false.     ! set .false. again by whoever overwrites the host state (restart read, DA, IAU)")
    A("   logical, save :: summary_pending = .false.")
    A("   real (kind=RKIND), save :: lbc_prev_to_end = -1.0_RKIND     ! regional runs: seconds to the end of the LBC interval at the last step")
    A("")
    A("   abstract interface")
    A("      subroutine halo_exchange_routine(domain, halo_group, ierr)")
    A("         use mpas_derived_types, only : domain_type")
    A("         type (domain_type), intent(inout) :: domain")
    A("         character(len=*), intent(in) :: halo_group")
    A("         integer, intent(out), optional :: ierr")
    A("      end subroutine halo_exchange_routine")
    A("   end interface")
    A("")
    A("contains")
    A("")
    A("   subroutine mpasb_check(ierr, what)")
    A("      integer(c_int), intent(in) :: ierr")
    A("      character(len=*), intent(in) :: what")
    A("      ! error convention of the ABI: 0 or a nonzero code + message; fatal = MPAS_LOG_CRIT (mpas_log.F:627-629)")
    A("      if (ierr /= 0) then")
    A("         call mpas_log_write('libmpasb: '//trim(what)//' failed: '//mpasb_c_to_f(mpasb_last_error(mpasb_h)), messageType=MPAS_LOG_CRIT)")
    A("      end if")
    A("   end subroutine mpasb_check")
    A("")
    # generic put/get helpers by rank and type
    for rk in (1, 2, 3):
        dims = ",".join(":" * 1 for _ in range(rk))
        A(f"   subroutine put_r{rk}(pool, key, devkey, lev, has_levels)")
        A("      type (mpas_pool_type), intent(in) :: pool")
        A("      character(len=*), intent(in) :: key, devkey")
        A("      integer, intent(in) :: lev")
        A("      logical, intent(in) :: has_levels")
        A(f"      real (kind=RKIND), dimension({dims}), pointer :: a")
        A("      if (has_levels) then")
        A("         call mpas_pool_get_array(pool, key, a, lev)")
        A("      else")
        A("         call mpas_pool_get_array(pool, key, a)")
        A("      end if")
        A("      call mpasb_check(mpasb_set_field(mpasb_h, trim(devkey)//c_null_char, int(lev, c_int), a, size(a, kind=c_long)), 'set_field '//devkey)")
        A(f"   end subroutine put_r{rk}")
        A("")
        A(f"   subroutine get_r{rk}(pool, key, devkey, lev, has_levels)")
        A("      type (mpas_pool_type), intent(in) :: pool")
        A("      character(len=*), intent(in) :: key, devkey")
        A("      integer, intent(in) :: lev")
        A("      logical, intent(in) :: has_levels")
        A(f"      real (kind=RKIND), dimension({dims}), pointer :: a")
        A("      if (has_levels) then")
        A("         call mpas_pool_get_array(pool, key, a, lev)")
        A("      else")
        A("         call mpas_pool_get_array(pool, key, a)")
        A("      end if")
        A("      call mpasb_check(mpasb_get_field(mpasb_h, trim(devkey)//c_null_char, int(lev, c_int), a, size(a, kind=c_long)), 'get_field '//devkey)")
        A(f"   end subroutine get_r{rk}")
        A("")
    for rk in (1, 2):
        dims = ",".join(":" for _ in range(rk))
        A(f"   subroutine put_i{rk}(pool, key)")
        A("      type (mpas_pool_type), intent(in) :: pool")
        A("      character(len=*), intent(in) :: key")
        A(f"      integer, dimension({dims}), pointer :: a")
        A("      call mpas_pool_get_array(pool, key, a)")
        A("      call mpasb_check(mpasb_set_field_int(mpasb_h, trim(key)//c_null_char, a, size(a, kind=c_long)), 'set_field_int '//key)")
        A(f"   end subroutine put_i{rk}")
        A("")

    def put(f, pool_var, key, lev=1, has_levels=False, ind="      "):
        if f["type"] == "INT":
            return f"{ind}call put_i{rank_of(f)}({pool_var}, '{key}')"
        return f"{ind}call put_r{rank_of(f)}({pool_var}, '{key}', '{f['name']}', {lev}, {'.true.' if has_levels else '.false.'})"

    def get(f, pool_var, key, lev=1, has_levels=False, ind="      "):
        return f"{ind}call get_r{rank_of(f)}({pool_var}, '{key}', '{f['name']}', {lev}, {'.true.' if has_levels else '.false.'})"

    # ---- exchange lists
    A("   ! field % sendList / recvList (mpas_field_types.inc:37-38; built by mpas_dmpar.F:1598-2158) of one cell, edge or vertex field,")
    A("   ! flattened per (neighbour, halo layer) as mpasb_set_halo_lists wants them: 1-based local indices, neighbours in ascending rank")
    A("   subroutine put_halo_lists(kind, sendList, recvList, nLayers, nprocs)")
    A("      integer, intent(in) :: kind, nLayers, nprocs")
    A("      type (mpas_multihalo_exchange_list), pointer :: sendList, recvList")
    A("      type (mpas_exchange_list), pointer :: e")
    A("      integer, allocatable :: nbr(:), nsend(:), nrecv(:), ssrc(:), rdst(:), slot(:)")
    A("      integer :: l, p, nn, ns, nr, i, os, or_")
    A("      allocate(slot(0:nprocs-1)); slot = 0")
    A("      do l = 1, nLayers                                   ! which ranks are neighbours at all")
    A("         e => sendList % halos(l) % exchList")
    A("         do while (associated(e)); slot(e % endPointID) = 1; e => e % next; end do")
    A("         e => recvList % halos(l) % exchList")
    A("         do while (associated(e)); slot(e % endPointID) = 1; e => e % next; end do")
    A("      end do")
    A("      nn = sum(slot); allocate(nbr(max(nn,1)), nsend(max(nn*nLayers,1)), nrecv(max(nn*nLayers,1)))")
    A("      nn = 0")
    A("      do p = 0, nprocs-1")
    A("         if (slot(p) == 1) then; nn = nn + 1; nbr(nn) = p; slot(p) = nn; end if")
    A("      end do")
    A("      nsend = 0; nrecv = 0; ns = 0; nr = 0")
    A("      do l = 1, nLayers")
    A("         e => sendList % halos(l) % exchList")
    A("         do while (associated(e))")
    A("            nsend((slot(e % endPointID)-1)*nLayers + l) = e % nList; ns = ns + e % nList; e => e % next")
    A("         end do")
    A("         e => recvList % halos(l) % exchList")
    A("         do while (associated(e))")
    A("            nrecv((slot(e % endPointID)-1)*nLayers + l) = e % nList; nr = nr + e % nList; e => e % next")
    A("         end do")
    A("      end do")
    A("      allocate(ssrc(max(ns,1)), rdst(max(nr,1)))")
    A("      os = 0; or_ = 0")
    A("      do p = 1, nn")
    A("         do l = 1, nLayers")
    A("            e => sendList % halos(l) % exchList")
    A("            do while (associated(e))")
    A("               if (slot(e % endPointID) == p) then")
    A("                  ! destList = position inside this neighbour's message (mpas_halo.F:1118-1121)")
    A("                  do i = 1, e % nList; ssrc(os + e % destList(i)) = e % srcList(i); end do")
    A("                  os = os + e % nList")
    A("               end if")
    A("               e => e % next")
    A("            end do")
    A("            e => recvList % halos(l) % exchList")
    A("            do while (associated(e))")
    A("               if (slot(e % endPointID) == p) then")
    A("                  do i = 1, e % nList; rdst(or_ + e % srcList(i)) = e % destList(i); end do     ! mpas_halo.F:1157-1162")
    A("                  or_ = or_ + e % nList")
    A("               end if")
    A("               e => e % next")
    A("            end do")
    A("         end do")
    A("      end do")
    A("      call mpasb_check(mpasb_set_halo_lists(mpasb_h, int(kind, c_int), int(nn, c_int), nbr, int(nLayers, c_int), nsend, ssrc, nrecv, rdst), 'set_halo_lists')")
    A("      deallocate(slot, nbr, nsend, nrecv, ssrc, rdst)")
    A("   end subroutine put_halo_lists")
    A("")
    # ---- init
    A("   ! TI:205, called from atm_core_init (mpas_atm_core.F:296) after atm_mpas_init_block: creates the device mirror of the block,")
    A("   ! uploads every mesh field of include/mpasb_fields.def by its pool key, the exchange lists and the communicator")
    A("   subroutine mpas_atm_dynamics_init(domain)")
    A("      type (domain_type), intent(inout) :: domain")
    A("      type (block_type), pointer :: block")
    A("      type (mpas_pool_type), pointer :: mesh, state, diag, tend, tend_physics")
    A("      type (mpasb_dims) :: dims")
    A("      type (mpasb_config) :: cfg")
    A("      integer, pointer :: ip")
    A("      real (kind=RKIND), pointer :: rp")
    A("      logical, pointer :: lp")
    A("      character (len=StrKIND), pointer :: sp")
    A("      type (field1DInteger), pointer :: idCell, idEdge, idVertex")
    A("      character(kind=c_char) :: uid(128), my_handles(128)")
    A("      character(kind=c_char), allocatable :: all_handles(:)")
    A("      integer :: local_gpu, n_gpus_per_node, mpi_ierr")
    A("      integer(c_long) :: nmax, nmax_global")
    A("      integer :: ok, ok_all")
    A("")
    A("      block => domain % blocklist          ! one block per rank == one block per GPU")
    A("      call mpas_pool_get_subpool(block % structs, 'mesh', mesh)")
    A("      call mpas_pool_get_subpool(block % structs, 'state', state)")
    A("      call mpas_pool_get_subpool(block % structs, 'diag', diag)")
    A("      call mpas_pool_get_subpool(block % structs, 'tend', tend)")
    A("      call mpas_pool_get_subpool(block % structs, 'tend_physics', tend_physics)")
    A("")
    header = open(os.path.join(ROOT, "include", "mpasb.h")).read()
    dim_src = {"num_scalars": "state", "index_qv": "state", "moist_start": "state", "moist_end": "state"}
    for ctype, n in struct_members(header, "mpasb_dims"):
        A(f"      call mpas_pool_get_dimension({dim_src.get(n, 'mesh')}, '{n}', ip); dims % {n} = ip")
    A("")
    for ctype, n in struct_members(header, "mpasb_config"):
        if n in ("cf1", "cf2", "cf3"):
            A(f"      call mpas_pool_get_array(mesh, '{n}', rp); cfg % {n} = rp")
        elif n == "sphere_radius":
            A("      call mpas_pool_get_config(mesh, 'sphere_radius', rp); cfg % sphere_radius = rp")
        elif n == "on_a_sphere":
            A("      call mpas_pool_get_config(mesh, 'on_a_sphere', lp); cfg % on_a_sphere = merge(1, 0, lp)")
        elif n == "config_horiz_mixing":
            A("      call mpas_pool_get_config(block % configs, 'config_horiz_mixing', sp)")
            A("      cfg % config_horiz_mixing = merge(0, 1, trim(sp) == '2d_smagorinsky')       ! 0 = 2d_smagorinsky, 1 = 2d_fixed (TI:5226, 5261)")
        elif n in ("config_split_dynamics_transport", "config_scalar_advection", "config_monotonic", "config_positive_definite",
                   "config_mix_full", "config_rayleigh_damp_u", "config_apply_lbcs", "config_print_global_minmax_vel"):
            A(f"      call mpas_pool_get_config(block % configs, '{n}', lp); cfg % {n} = merge(1, 0, lp)")
        elif ctype == "int":
            A(f"      call mpas_pool_get_config(block % configs, '{n}', ip); cfg % {n} = ip")
        else:
            A(f"      call mpas_pool_get_config(block % configs, '{n}', rp); cfg % {n} = rp")
    A("")
    A("      n_gpus_per_node = 8")
    A("      local_gpu = mod(domain % dminfo % my_proc_id, n_gpus_per_node)")
    A("      call mpasb_check(mpasb_create(dims, cfg, int(local_gpu, c_int), mpasb_h), 'create')")
    A("")
    A("      ! mesh fields (what TI:291-459 copies to the device in the OpenACC build, plus coeffs_reconstruct / latCell / lonCell for")
    A("      ! the step's trailing mpas_reconstruct, TI:1606)")
    for f in mesh:
        A(put(f, "mesh", f["name"]))
    A("")
    A("      ! exchange lists of the three element kinds: 2 cell layers, 3 edge and vertex layers (mpas_bootstrapping.F)")
    A("      call mpas_pool_get_field(mesh, 'indexToCellID', idCell)")
    A("      call mpas_pool_get_field(mesh, 'indexToEdgeID', idEdge)")
    A("      call mpas_pool_get_field(mesh, 'indexToVertexID', idVertex)")
    A("      call put_halo_lists(0, idCell % sendList, idCell % recvList, 2, domain % dminfo % nprocs)")
    A("      call put_halo_lists(1, idEdge % sendList, idEdge % recvList, 3, domain % dminfo % nprocs)")
    A("      call put_halo_lists(2, idVertex % sendList, idVertex % recvList, 3, domain % dminfo % nprocs)")
    A("      if (domain % dminfo % nprocs > 1) then")
    A("         if (domain % dminfo % my_proc_id == 0) call mpasb_check(mpasb_get_nccl_unique_id(uid), 'get_nccl_unique_id')")
    A("         call mpas_dmpar_bcast_chars(domain % dminfo, 128, uid)")
    A("         call mpasb_check(mpasb_comm_init(mpasb_h, int(domain % dminfo % my_proc_id, c_int), int(domain % dminfo % nprocs, c_int), uid), 'comm_init')")
    A("         ! ranks of one node: exchanges by direct NVLink stores (CUDA IPC); all ranks switch together or none does")
    A("         nmax = mpasb_p2p_max_message(mpasb_h)")
    A("         call mpas_dmpar_max_int(domain % dminfo, int(nmax), ok); nmax_global = ok")
    A("         ok = merge(1, 0, nmax >= 0 .and. domain % dminfo % nprocs <= 9)")
    A("         if (ok == 1) ok = merge(1, 0, mpasb_p2p_prepare(mpasb_h, nmax_global, my_handles) == 0)")
    A("         call mpas_dmpar_min_int(domain % dminfo, ok, ok_all)")
    A("         if (ok_all == 1) then")
    A("            allocate(all_handles(128 * domain % dminfo % nprocs))")
    A("            call MPI_Allgather(my_handles, 128, MPI_BYTE, all_handles, 128, MPI_BYTE, domain % dminfo % comm, mpi_ierr)")
    A("            ok = merge(1, 0, mpasb_p2p_open(mpasb_h, all_handles) == 0)")
    A("            call mpas_dmpar_min_int(domain % dminfo, ok, ok_all)")
    A("            if (ok_all == 1) call mpasb_check(mpasb_p2p_enable(mpasb_h, 1_c_int), 'p2p_enable')")
    A("            deallocate(all_handles)")
    A("         end if")
    A("      end if")
    A("      state_on_device = .false.")
    A("   end subroutine mpas_atm_dynamics_init")
    A("")
    A("   ! TI:479, called from atm_core_finalize (mpas_atm_core.F:1054)")
    A("   subroutine mpas_atm_dynamics_finalize(domain)")
    A("      type (domain_type), intent(inout) :: domain")
    A("      integer(c_int) :: ierr")
    A("      ierr = mpasb_destroy(mpasb_h)")
    A("      mpasb_h = c_null_ptr")
    A("   end subroutine mpas_atm_dynamics_finalize")
    A("")
    # ---- upload state
    A("   ! host pools -> device: prognostic state (time level 1), coupled diagnostics, tendencies the physics left in the pools")
    A("   subroutine mpasb_upload_state(block)")
    A("      type (block_type), intent(inout) :: block")
    A("      type (mpas_pool_type), pointer :: state, diag, tend, tend_physics")
    A("      call mpas_pool_get_subpool(block % structs, 'state', state)")
    A("      call mpas_pool_get_subpool(block % structs, 'diag', diag)")
    A("      call mpas_pool_get_subpool(block % structs, 'tend', tend)")
    A("      call mpas_pool_get_subpool(block % structs, 'tend_physics', tend_physics)")
    for n in STATE:
        A(put(byname[n], "state", n, 1, True))
    for n in DIAG_IN:
        A(put(byname[n], "diag", n))
    A(put(byname["rt_diabatic_tend"], "tend", "rt_diabatic_tend"))
    A("   end subroutine mpasb_upload_state")
    A("")
    A("   ! regional runs: the driving fields of the current LBC interval, var_struct 'lbc' (time level 1 = tendency, 2 = state at the")
    A("   ! end of the interval, mpas_atm_boundaries.F:78-339), re-sent whenever mpas_atm_update_bdy_tend has read a new LBC time")
    A("   subroutine mpasb_upload_lbc(block)")
    A("      type (block_type), intent(inout) :: block")
    A("      type (mpas_pool_type), pointer :: lbc")
    A("      call mpas_pool_get_subpool(block % structs, 'lbc', lbc)")
    for n in LBC:
        for lev in (1, 2):
            A(put(byname[n], "lbc", n, lev, True))
    A("   end subroutine mpasb_upload_lbc")
    A("")
    A("   ! device -> host pools on output / restart alarms (mpas_atm_core.F:812-890), before the stream write")
    A("   subroutine mpasb_download_for_output(block)")
    A("      type (block_type), intent(inout) :: block")
    A("      type (mpas_pool_type), pointer :: state, diag, tend_physics")
    A("      call mpas_pool_get_subpool(block % structs, 'state', state)")
    A("      call mpas_pool_get_subpool(block % structs, 'diag', diag)")
    A("      call mpas_pool_get_subpool(block % structs, 'tend_physics', tend_physics)")
    A("      call mpasb_check(mpasb_compute_output_diagnostics(mpasb_h, 1_c_int), 'compute_output_diagnostics')      ! mpas_atm_core.F:901-950")
    for n in STATE:
        A(get(byname[n], "state", n, 1, True))
    for n in DIAG_OUT:
        A(get(byname[n], "diag", n))
    A(get(byname["rthdynten"], "tend_physics", "rthdynten"))
    A("   end subroutine mpasb_download_for_output")
    A("")
    # ---- timestep
    A("   ! TI:739: same signature; exchange_halo_group is accepted and unused (halo exchanges run on the GPU inside mpasb_step)")
    A("   subroutine atm_timestep(domain, dt, nowTime, itimestep, exchange_halo_group)")
    A("      type (domain_type), intent(inout) :: domain")
    A("      real (kind=RKIND), intent(in) :: dt")
    A("      type (MPAS_Time_type), intent(in) :: nowTime")
    A("      integer, intent(in) :: itimestep")
    A("      procedure (halo_exchange_routine) :: exchange_halo_group")
    A("      type (block_type), pointer :: block")
    A("      type (mpas_pool_type), pointer :: state")
    A("      character (len=StrKIND), pointer :: config_time_integration, xtime")
    A("      logical, pointer :: config_print_global_minmax_vel, config_print_global_minmax_sca, config_apply_lbcs")
    A("      integer, pointer :: num_scalars")
    A("      type (MPAS_TimeInterval_type) :: lbc_interval")
    A("      integer :: dd_intv, s_intv, sn_intv, sd_intv")
    A("      real (kind=RKIND) :: lbc_to_end")
    A("      real (kind=mpasb_real), allocatable :: mm(:)")
    A("      real (kind=RKIND) :: gmin, gmax")
    A("      integer(c_long) :: nan_count(2)")
    A("      type (MPAS_Time_type) :: currTime")
    A("      type (MPAS_TimeInterval_type) :: dtInterval")
    A("      integer :: ierr")
    A("")
    A("      block => domain % blocklist")
    A("      call mpas_pool_get_config(block % configs, 'config_time_integration', config_time_integration)")
    A("      if (trim(config_time_integration) /= 'SRK3') then")
    A("         call mpas_log_write('Unknown time integration option '//trim(config_time_integration), messageType=MPAS_LOG_CRIT)      ! TI:786-790")
    A("      end if")
    A("      call mpas_pool_get_subpool(block % structs, 'state', state)")
    A("      call mpas_pool_get_dimension(state, 'num_scalars', num_scalars)")
    A("      call mpas_pool_get_config(block % configs, 'config_print_global_minmax_vel', config_print_global_minmax_vel)")
    A("      call mpas_pool_get_config(block % configs, 'config_print_global_minmax_sca', config_print_global_minmax_sca)")
    A("")
    A("      if (.not. state_on_device) then          ! first call, or the host state was rewritten (restart read, DA, IAU)")
    A("         call mpasb_upload_state(block)")
    A("         state_on_device = .true.")
    A("      end if")
    A("")
    A("      call mpas_pool_get_config(block % configs, 'config_apply_lbcs', config_apply_lbcs)")
    A("      if (config_apply_lbcs) then              ! regional run (TI:773): what mpas_atm_get_bdy_state derives from the clock,")
    A("         ! mpas_atm_boundaries.F:497-503.  LBC_intv_end is private to mpas_atm_boundaries: mpas_atm_bdy_interval_end() is the one")
    A("         ! accessor a maintainer adds there (`function mpas_atm_bdy_interval_end() result(t); t = LBC_intv_end; end function`)")
    A("         lbc_interval = mpas_atm_bdy_interval_end() - nowTime")
    A("         call mpas_get_timeInterval(interval=lbc_interval, DD=dd_intv, S=s_intv, S_n=sn_intv, S_d=sd_intv, ierr=ierr)")
    A("         lbc_to_end = 86400.0_RKIND * real(dd_intv, kind=RKIND) + real(s_intv, kind=RKIND) &")
    A("                      + (real(sn_intv, kind=RKIND) / real(sd_intv, kind=RKIND))")
    A("         if (lbc_to_end > lbc_prev_to_end) call mpasb_upload_lbc(block)     ! a new LBC interval began (mpas_atm_core.F:724, 754)")
    A("         lbc_prev_to_end = lbc_to_end")
    A("         call mpasb_check(mpasb_set_lbc_time(mpasb_h, real(lbc_to_end, mpasb_real)), 'set_lbc_time')")
    A("      end if")
    A("")
    A("      call mpasb_check(mpasb_step(mpasb_h, real(dt, mpasb_real), int(itimestep, c_int)), 'step')      ! atm_srk3, TI:803-1725")
    A("")
    A("      if (config_print_global_minmax_vel .or. config_print_global_minmax_sca) then              ! summarize_timestep, TI:7914-8357")
    A("         allocate(mm(2*(2+num_scalars)))")
    A("         call mpasb_check(mpasb_summarize_timestep_async(mpasb_h), 'summarize_timestep_async')")
    A("         call mpasb_check(mpasb_summarize_timestep_fetch(mpasb_h, mm, int(size(mm), c_long), nan_count), 'summarize_timestep_fetch')")
    A("         if (nan_count(1) > 0) call mpas_log_write('NaN detected in ''w'' field.', messageType=MPAS_LOG_CRIT)      ! TI:8268")
    A("         if (nan_count(2) > 0) call mpas_log_write('NaN detected in ''u'' field.', messageType=MPAS_LOG_CRIT)      ! TI:8280")
    A("         call mpas_dmpar_min_real(domain % dminfo, real(mm(1), RKIND), gmin)")
    A("         call mpas_dmpar_max_real(domain % dminfo, real(mm(2), RKIND), gmax)")
    A("         call mpas_log_write('global min, max w $r $r', realArgs=(/gmin, gmax/))                  ! TI:8304")
    A("         call mpas_dmpar_min_real(domain % dminfo, real(mm(3), RKIND), gmin)")
    A("         call mpas_dmpar_max_real(domain % dminfo, real(mm(4), RKIND), gmax)")
    A("         call mpas_log_write('global min, max u $r $r', realArgs=(/gmin, gmax/))                  ! TI:8319")
    A("         deallocate(mm)")
    A("      end if")
    A("")
    A("      ! TI:792-798: time stamp of the new state")
    A("      call mpas_pool_get_array(state, 'xtime', xtime, 2)")
    A("      call mpas_set_timeInterval(dtInterval, dt=dt)")
    A("      currTime = nowTime + dtInterval")
    A("      call mpas_get_time(currTime, dateTimeString=xtime)")
    A("")
    A("      ! mpas_atm_core.F:808 then calls mpas_pool_shift_time_levels(state) on the host pools; the device mirror rotates with it")
    A("      call mpasb_check(mpasb_shift_time_levels(mpasb_h), 'shift_time_levels')")
    A("   end subroutine atm_timestep")
    A("")
    A("end module atm_time_integration_b200")
    A("")
    return "\n".join(L)


def MESH_KEYS(F):
    txt = open(os.path.join(ROOT, "include", "mpasb_fields.def")).read()
    start = txt.index("/* ---- mesh, real ---- */")
    return set(re.findall(r"^F\((\w+),", txt[start:], re.M))


def wrap(text, width=124):
    """Fortran continuation lines: break at a comma or an operator-free blank outside strings and comments."""
    out = []
    for line in text.split("\n"):
        code = line
        while len(code) > width and not code.lstrip().startswith(("!", "#")):
            bang = code.find(" ! ")
            limit = width if bang < 0 or bang > width else bang
            cut, q = -1, None
            for j, ch in enumerate(code[:limit]):
                if q:
                    if ch == q:
                        q = None
                elif ch in "'\"":
                    q = ch
                elif ch == "," and j > 40 and code[j - 6:j] != "bind(C":
                    cut = j
            if cut < 0:
                break
            out.append(code[:cut + 1] + " &")
            code = " " * (len(line) - len(line.lstrip()) + 6) + code[cut + 1:].lstrip()
        out.append(code)
    return "\n".join(out)


def main():
    os.makedirs(os.path.join(ROOT, "fortran"), exist_ok=True)
    out = {"mpasb_binding.F90": wrap(binding()), "mpas_atm_dynamics_b200.F": wrap(dynamics())}
    for name, text in out.items():
        with open(os.path.join(ROOT, "fortran", name), "w") as f:
            f.write(text)
    return out


if __name__ == "__main__":
    main()
