"""Is NVLink-switch multicast (the substrate of multimem.ld_reduce / multimem.st, "NVLS") usable on this box?  One process, all
visible GPUs: device attribute, multicast granularity, cuMulticastCreate + AddDevice + a bound and mapped 2 MiB page per device.
Prints one JSON line; every driver error is reported, nothing is assumed."""
import json

from cuda.bindings import driver as cu


def ck(r, what, out):
    err = r[0]
    if int(err) != 0:
        out.setdefault("errors", []).append(f"{what}: {cu.cuGetErrorName(err)[1].decode() if cu.cuGetErrorName(err)[0] == 0 else int(err)}")
        raise RuntimeError(what)
    return r[1] if len(r) == 2 else r[1:]


def main():
    out = {}
    try:
        ck(cu.cuInit(0), "cuInit", out)
        n = ck(cu.cuDeviceGetCount(), "cuDeviceGetCount", out)
        out["devices"] = n
        devs = [ck(cu.cuDeviceGet(i), "cuDeviceGet", out) for i in range(n)]
        out["multicast_supported"] = [ck(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d), "attr", out) for d in devs]
        out["fabric_handle_supported"] = [ck(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, d), "attr", out) for d in devs]
        out["posix_fd_handle_supported"] = [ck(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, d), "attr", out) for d in devs]
        if n < 2 or not all(out["multicast_supported"]):
            print(json.dumps(out))
            return
        ctxs = [ck(cu.cuDevicePrimaryCtxRetain(d), "ctx", out) for d in devs]
        ck(cu.cuCtxSetCurrent(ctxs[0]), "setctx", out)
        prop = cu.CUmulticastObjectProp()
        prop.numDevices = n
        prop.size = 2 << 20
        prop.handleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        gran = ck(cu.cuMulticastGetGranularity(prop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED), "granularity", out)
        out["multicast_granularity"] = int(gran)
        prop.size = max(int(gran), 2 << 20)
        mc = ck(cu.cuMulticastCreate(prop), "cuMulticastCreate", out)
        for d in devs:
            ck(cu.cuMulticastAddDevice(mc, d), "cuMulticastAddDevice", out)
        for i, d in enumerate(devs):
            ap = cu.CUmemAllocationProp()
            ap.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
            ap.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
            ap.location.id = i
            ap.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
            h = ck(cu.cuMemCreate(prop.size, ap, 0), "cuMemCreate", out)
            ck(cu.cuMulticastBindMem(mc, 0, h, 0, prop.size, 0), "cuMulticastBindMem", out)
        va = ck(cu.cuMemAddressReserve(prop.size, int(gran), 0, 0), "reserve", out)
        ck(cu.cuMemMap(va, prop.size, 0, mc, 0), "cuMemMap(multicast)", out)
        out["multicast_object_bound_and_mapped"] = True
    except RuntimeError:
        pass
    print(json.dumps(out))


if __name__ == "__main__":
    main()
