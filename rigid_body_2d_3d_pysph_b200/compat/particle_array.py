"""Host-side particle container with the PySPH ``ParticleArray`` surface.

The reference scripts build their scenes through ``get_particle_array`` and
then poke at the result as if every property were a NumPy array
(``pa.x[:] += ...``, ``max(pa.body_id)``, ``pa.add_property(name, type=,
data=, stride=)``, ``pa.add_constant``, ``pa.remove_particles(idx)``; see
/root/reference/code/benchmark_5_steady_cubes_on_a_wall_3d.py:298-360 and
/root/reference/code/stack_of_cylinders.py:133-243).  PySPH itself is not
available, so this module restates the part of that surface the rigid-body
path touches.

Storage is plain NumPy on the host.  When a device scene is bound (see
``rigid_body_2d_3d_pysph_b200.device.DeviceScene``) the array follows a lazy
coherence protocol: reading ``pa.<prop>`` pulls the property from the GPU if
the device copy is newer and marks it "host touched", so that it is pushed
back before the next step.  No arithmetic of the hot path runs here.
"""
import zlib

import numpy as np

_TYPE_MAP = {
    'double': np.float64,
    'float': np.float64,
    'int': np.int32,
    'long': np.int64,
    'unsigned int': np.uint32,
}

_DEFAULT_PROPS = ['x', 'y', 'z', 'u', 'v', 'w', 'm', 'h', 'rho', 'p',
                  'au', 'av', 'aw']
_DEFAULT_INT_PROPS = {'gid': 'unsigned int', 'pid': 'int', 'tag': 'int'}


class ParticleArray(object):
    """NumPy-backed restatement of PySPH's ParticleArray (App. B)."""

    def __init__(self, name='', default_particle_tag=0, constants=None,
                 backend=None, **props):
        d = self.__dict__
        d['name'] = name
        d['properties'] = {}
        d['constants'] = {}
        d['stride'] = {}
        d['property_types'] = {}
        d['output_property_arrays'] = []
        d['default_values'] = {}
        d['num_real_particles'] = 0
        d['backend'] = backend
        d['time'] = 0.0
        # lazy device coherence (filled by DeviceScene.bind)
        d['_device'] = None
        d['_device_newer'] = set()
        d['_host_touched'] = set()
        d['_host_read'] = {}
        self._n = 0
        if props:
            self._initialize(**props)
        if constants:
            for key, val in constants.items():
                self.add_constant(key, val)

    # ------------------------------------------------------------------
    def _initialize(self, **props):
        n = 0
        conv = {}
        for key, val in props.items():
            if isinstance(val, dict):
                data = val.get('data', None)
                spec = dict(val)
            else:
                data = val
                spec = {}
            if data is not None:
                arr = np.ravel(np.asarray(data))
                stride = spec.get('stride', 1)
                if arr.size > 1 or np.ndim(data) > 0:
                    n = max(n, arr.size // stride)
            conv[key] = (data, spec)
        self.__dict__['_n'] = n
        for key, (data, spec) in conv.items():
            self.add_property(key, type=spec.get('type', 'double'),
                              default=spec.get('default', None), data=data,
                              stride=spec.get('stride', 1))
        self.__dict__['num_real_particles'] = n

    # ------------------------------------------------------------------
    def get_number_of_particles(self, real=False):
        return self._n

    def __len__(self):
        return self._n

    # ------------------------------------------------------------------
    def add_property(self, name, type='double', default=None, data=None,
                     stride=1):
        dtype = _TYPE_MAP[type]
        stride = int(stride)
        n = self._n
        if default is None:
            default = 0
        if data is None:
            arr = np.full(n * stride, default, dtype=dtype)
        else:
            src = np.asarray(data)
            if src.ndim == 0:
                arr = np.full(n * stride, src.item(), dtype=dtype)
            else:
                src = np.ravel(src)
                if n == 0 and not self.properties:
                    n = src.size // stride
                    self.__dict__['_n'] = n
                    self.__dict__['num_real_particles'] = n
                if src.size != n * stride:
                    raise ValueError(
                        'property %s: size %d != %d particles x stride %d' %
                        (name, src.size, n, stride))
                arr = np.array(src, dtype=dtype, copy=True)
        if name in self.properties and \
                self.properties[name].dtype == arr.dtype and \
                self.properties[name].size == arr.size:
            # PySPH keeps the existing array object and overwrites when data
            # is given; views handed out earlier stay valid.
            if data is not None:
                self.properties[name][:] = arr
            self.stride[name] = stride
            self._touch(name)
            return
        self.properties[name] = arr
        self.stride[name] = stride
        self.property_types[name] = type
        self.default_values[name] = default
        self._touch(name)

    def add_constant(self, name, data):
        if isinstance(data, (int, float, np.integer, np.floating)):
            data = [data]
        arr = np.ravel(np.array(data, copy=True))
        if arr.dtype.kind == 'f':
            arr = arr.astype(np.float64)
        elif arr.dtype.kind in 'iu':
            arr = arr.astype(np.int64)
        self.constants[name] = arr
        self._touch(name)

    def remove_property(self, name):
        self.properties.pop(name, None)
        self.stride.pop(name, None)
        if name in self.output_property_arrays:
            self.output_property_arrays.remove(name)

    # ------------------------------------------------------------------
    def _touch(self, name):
        self.__dict__['_host_touched'].add(name)
        self.__dict__['_device_newer'].discard(name)

    def touch(self, *names):
        """Declare properties / constants modified in place by host code."""
        for name in names:
            self._touch(name)

    def _note_read(self, name, arr):
        """A read hands out the live array, which the caller may modify in
        place (``pa.x[:] += 0.25``) without this object ever hearing of it.
        While a device scene is bound, remember a checksum of what was handed
        out (once per name between two steps); ``modified_since_read`` then
        tells a read from a write, so that a callback that only LOOKS at
        ``pa.m`` or ``pa.x`` does not force an upload, a static-table rebuild
        or a neighbour-list rebuild at the next step."""
        d = self.__dict__
        if d['_device'] is None:
            d['_host_touched'].add(name)
            return
        if name in d['_host_touched'] or name in d['_host_read']:
            return
        d['_host_read'][name] = zlib.crc32(memoryview(arr).cast('B'))

    def modified_since_read(self):
        """Names handed out since the last call whose content has changed."""
        d = self.__dict__
        out = set()
        for name, crc in d['_host_read'].items():
            arr = d['properties'].get(name)
            if arr is None:
                arr = d['constants'].get(name)
            if arr is None or zlib.crc32(memoryview(arr).cast('B')) != crc:
                out.add(name)
        d['_host_read'].clear()
        return out

    def _pull(self, name):
        d = self.__dict__
        if name in d['_device_newer'] and d['_device'] is not None:
            d['_device'].pull(self, name)
            d['_device_newer'].discard(name)

    def __getattr__(self, name):
        d = self.__dict__
        props = d.get('properties')
        if props is not None and name in props:
            self._pull(name)
            self._note_read(name, props[name])
            return props[name]
        consts = d.get('constants')
        if consts is not None and name in consts:
            self._pull(name)
            self._note_read(name, consts[name])
            return consts[name]
        raise AttributeError("ParticleArray '%s' has no property or constant "
                             "'%s'" % (d.get('name'), name))

    def __setattr__(self, name, value):
        d = self.__dict__
        if name in d.get('properties', ()):
            arr = d['properties'][name]
            if value is not arr:
                arr[:] = value
            self._touch(name)
        elif name in d.get('constants', ()):
            arr = d['constants'][name]
            if value is not arr:
                arr[:] = value
            self._touch(name)
        else:
            d[name] = value

    # ------------------------------------------------------------------
    def get(self, *names, **kw):
        out = [getattr(self, n) for n in names]
        return out[0] if len(out) == 1 else out

    def set(self, **props):
        for key, val in props.items():
            setattr(self, key, val)

    def get_property_arrays(self, all=True, only_real=True):
        names = list(self.properties) if all or \
            not self.output_property_arrays else self.output_property_arrays
        return dict((n, getattr(self, n)) for n in names)

    def set_output_arrays(self, props):
        for p in props:
            if p not in self.properties:
                raise ValueError('%s is not a property of %s' %
                                 (p, self.name))
        self.__dict__['output_property_arrays'] = list(props)

    def add_output_arrays(self, props):
        cur = self.output_property_arrays
        for p in props:
            if p not in self.properties:
                raise ValueError('%s is not a property of %s' %
                                 (p, self.name))
            if p not in cur:
                cur.append(p)

    # ------------------------------------------------------------------
    def remove_particles(self, indices, align=True):
        """Swap-with-last removal, processed from the largest index down.

        [upstream] cyarray ``BaseArray.remove``: indices are sorted, then for
        each one (largest first) the last live element is copied into the
        hole and the length shrinks by one.  Particle order is therefore not
        preserved and a duplicated index removes one extra particle.  The
        reference relies on it at stack_of_cylinders.py:211-233.
        """
        idx = np.sort(np.ravel(np.asarray(indices, dtype=np.int64)))
        if idx.size == 0:
            return
        n = self._n
        if idx.size > n:
            raise ValueError('Number of particles to be removed is greater '
                             'than number of particles in array')
        for name in list(self.properties):
            self._pull(name)
        length = n
        moves = []
        for i in idx[::-1]:
            if i < length:
                moves.append((int(i), length - 1))
                length -= 1
        for name, arr in self.properties.items():
            s = self.stride[name]
            for dst, src in moves:
                arr[dst * s:(dst + 1) * s] = arr[src * s:(src + 1) * s]
            self.properties[name] = arr[:length * s].copy()
            self._touch(name)
        self.__dict__['_n'] = length
        self.__dict__['num_real_particles'] = length

    def extend(self, num_particles):
        if num_particles <= 0:
            return
        for name, arr in self.properties.items():
            self._pull(name)
            s = self.stride[name]
            extra = np.full(num_particles * s, self.default_values.get(name, 0),
                            dtype=arr.dtype)
            self.properties[name] = np.concatenate([arr, extra])
            self._touch(name)
        self.__dict__['_n'] = self._n + num_particles
        self.__dict__['num_real_particles'] = self._n

    def add_particles(self, **props):
        if not props:
            return
        first = np.ravel(np.asarray(next(iter(props.values()))))
        key0 = next(iter(props))
        num = first.size // self.stride.get(key0, 1)
        old = self._n
        self.extend(num)
        for key, val in props.items():
            s = self.stride[key]
            self.properties[key][old * s:] = np.ravel(np.asarray(val))

    def extract_particles(self, indices, dest_array=None, align=True,
                          props=None):
        idx = np.ravel(np.asarray(indices, dtype=np.int64))
        out = dest_array if dest_array is not None else \
            ParticleArray(name=self.name)
        names = props if props is not None else list(self.properties)
        out.__dict__['_n'] = idx.size
        for name in names:
            s = self.stride[name]
            src = getattr(self, name).reshape(-1, s)[idx].ravel()
            out.add_property(name, type=self.property_types[name], data=src,
                             stride=s)
        for name, val in self.constants.items():
            out.add_constant(name, val)
        out.set_output_arrays([p for p in self.output_property_arrays
                               if p in out.properties])
        return out

    def align_particles(self):
        pass

    def set_name(self, name):
        self.__dict__['name'] = name

    def get_carray(self, name):
        return getattr(self, name)

    def __repr__(self):
        return "<ParticleArray '%s' n=%d props=%d consts=%d>" % (
            self.name, self._n, len(self.properties), len(self.constants))


def get_particle_array(additional_props=None, constants=None, backend=None,
                       **props):
    """[upstream] ``pysph.base.utils.get_particle_array``.

    Default properties x y z u v w m h rho p au av aw (double), gid (unsigned
    int), pid, tag (int); scalars broadcast to the longest array given.
    Used at /root/reference/code/benchmark_5_steady_cubes_on_a_wall_3d.py:298.
    """
    name = props.pop('name', '')
    nprops = {}
    n = 0
    for key, val in props.items():
        arr = np.asarray(val)
        if arr.ndim > 0:
            n = max(n, arr.size)
    for key in _DEFAULT_PROPS:
        if key not in props:
            nprops[key] = {'data': np.zeros(n), 'type': 'double'}
    for key, val in props.items():
        arr = np.asarray(val)
        if key in _DEFAULT_INT_PROPS:
            typ = _DEFAULT_INT_PROPS[key]
        else:
            typ = 'double'
        if arr.ndim == 0:
            data = np.full(n, arr.item())
        else:
            data = np.ravel(arr)
        nprops[key] = {'data': data, 'type': typ}
    for key, typ in _DEFAULT_INT_PROPS.items():
        if key not in nprops:
            nprops[key] = {'data': np.zeros(n, dtype=np.int64), 'type': typ}
    if additional_props:
        for key in additional_props:
            if key not in nprops:
                nprops[key] = {'data': np.zeros(n), 'type': 'double'}
    pa = ParticleArray(name=name, constants=constants, backend=backend,
                       **nprops)
    pa.set_output_arrays(['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'm', 'h',
                          'pid', 'gid', 'tag', 'p'])
    return pa
