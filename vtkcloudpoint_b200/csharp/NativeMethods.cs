// NativeMethods.cs -- P/Invoke declarations for libvpc (include/vpc.h), in the style of the only P/Invoke
// precedent in the reference tree (vtkPointCloud/BaseClass/FileMap.cs:73-130, [DllImport("kernel32.dll")]).
// Source only: the build image has no .NET toolchain, so this file is reviewed, not compiled (DESIGN.md).
using System;
using System.Runtime.InteropServices;

namespace vtkPointCloud
{
    internal static class NativeMethods
    {
        private const string Lib = "vpc";   // vpc.dll on Windows, libvpc.so elsewhere (probing path: app.config:5-7)

        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_create(out IntPtr ctx, int[] deviceIds, int nDevices);

        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern void vpc_destroy(IntPtr ctx);

        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern IntPtr vpc_last_error(IntPtr ctx);

        // DBImproved.dbscan (BaseClass/DBImproved.cs:91-114)
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_dbscan_l1_2d(IntPtr ctx, double[] mx, double[] my, long n, double eps, int minPts,
            int firstClusterId, [Out] int[] clusterId, [Out] byte[] isKey, [Out] byte[] isClassed, out int clusterAmount);

        // every StartCode work item (FrmMain.cs:2782-2794) in one call
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_dbscan_l1_2d_cells(IntPtr ctx, double[] mx, double[] my, long n, long[] cellOffsets, int nCells,
            double eps, int minPts, [Out] int[] clusterId, [Out] byte[] isKey, [Out] byte[] isClassed, [Out] int[] clusterAmountPerCell);

        // getClusterFromMotor -> DoWork3 -> CompleteWork3 up to the renumbered labels (FrmMain.cs:1214-1291, 1340-1361, 1432-1520), one call
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_dbscan_blocked_ref(IntPtr ctx, double[] mx, double[] my, long n, double eps, int minPts, int ptsInCell,
            [Out] int[] clusterId, out int clusterSum, out int delSum, out int rows, out int cols, out long nUnassigned);

        // ICP.FindClosestPointSet (BaseClass/ICP.cs:224-250); xyz arrays are planar x[0..k) y[0..k) z[0..k)
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_closest_point_set(IntPtr ctx, double[] modelXyz, long m, double[] dataXyz, long n,
            [Out] int[] order, [Out] double[] sqdist);

        // ICP.go_hell_ICP (BaseClass/ICP.cs:18-181)
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_icp_rigid(IntPtr ctx, double[] modelXyz, long m, double[] dataXyz, long n, double e,
            int maxIters, [In, Out] double[] R, [In, Out] double[] T, out int itersDone, out double sseLast, [Out] int[] orderLast);

        // MainForm.RecorrectMatchingPtsByDistance's search (FrmMain.cs:3588-3618)
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_match_within(IntPtr ctx, double[] truthXyz, long m, double[] centersXyz, long n, double matchDistance,
            [Out] int[] matchedId, [Out] double[] dist);

        // Tools.GetClusList + Tools.getCircles(3-D) + getCircles(2-D): the statistics block of CompleteWork3 (FrmMain.cs:1521-1540)
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_cluster_stats(IntPtr ctx, int[] clusterId, long n, int nClusters, double[] xyz, double[] mx, double[] my,
            [Out] double[] means5, [Out] int[] counts, [Out] double[] circle3d, [Out] int[] status3d, [Out] double[] circle2d, [Out] int[] status2d);

        // the LINQ query of MainForm.refreshClusList (FrmMain.cs:3446-3467)
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_nearest_truth_2d(IntPtr ctx, double[] truthX, double[] truthY, int[] truthId, long m, double[] px, double[] py,
            long n, double radius, [Out] int[] id);

        // the import loop (FrmMain.cs:975-1068): text rows -> motor_x, motor_y, Distance -> gate -> XYZ -> duplicate removal
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_ingest_text(IntPtr ctx, byte[] text, long len, double xAngle, double yAngle, int xdir, int ydir, int removeDuplicates,
            long rowCap, [Out] double[] mx, [Out] double[] my, [Out] double[] dist, [Out] double[] xyz, [Out] byte[] keep, [Out] byte[] rowStatus,
            out long nRows, out long nKept, out long nDuplicates);

        // the same call with the merge bookkeeping: shared-point count, clusForMerge (FrmMain.cs:1517-1520) as point indices + ids
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_dbscan_blocked_ref_ex(IntPtr ctx, double[] mx, double[] my, long n, double eps, int minPts, int ptsInCell,
            [Out] int[] clusterId, out int clusterSum, out int delSum, out int rows, out int cols, out long nUnassigned, out long nShared,
            [Out] long[] mergeOrder, [Out] int[] mergeCid, out long nMerge, out int clusterSumCells);

        // Clustering.MergeBtn_Click's chain: Tools.GetClusList -> MergeIDByDistance -> refreshCensAndClusByDictionary (Tools.cs:162-195, 580-621, 521-572)
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_merge_ids_by_distance(IntPtr ctx, int[] mergeCid, double[] xyz, double[] mx, double[] my, long k, int clusterAmount,
            double thre, [Out] int[] newCid, out int newAmount, [Out] int[] dictFrom, [Out] int[] dictTo, out int nDict, [Out] double[] centers5,
            [Out] int[] centerIds, out int nCenters, [Out] double[] newCenters5);

        // page-locked host memory: arrays allocated / registered here are copied at the PCIe rate without staging (include/vpc.h, "Host memory")
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_host_alloc(out IntPtr p, long bytes);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern void vpc_host_free(IntPtr p);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_host_register(IntPtr p, long bytes);
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl)]
        internal static extern int vpc_host_unregister(IntPtr p);
        // IntPtr overload of the clustering call for page-locked buffers
        [DllImport(Lib, CallingConvention = CallingConvention.Cdecl, EntryPoint = "vpc_dbscan_l1_2d")]
        internal static extern int vpc_dbscan_l1_2d_ptr(IntPtr ctx, IntPtr mx, IntPtr my, long n, double eps, int minPts,
            int firstClusterId, IntPtr clusterId, IntPtr isKey, IntPtr isClassed, out int clusterAmount);

        internal static void Check(IntPtr ctx, int rc)
        {
            if (rc != 0)   // the reference signals errors with MException (Matrix.cs:710-715)
                throw new MException("vpc error " + rc + ": " + Marshal.PtrToStringAnsi(vpc_last_error(ctx)));
        }
    }
}
